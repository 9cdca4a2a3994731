"""CPU oracle for the DiffAb denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline.  The product
package (``diffab-pytorch_b200/``) never imports it and has no CPU fallback.

Contents
--------
* ``so3``        restatement of ``diffab_pytorch/so3.py`` (maps, IGSO(3) table, sampler with injected noise)
* ``diffusion``  restatement of ``diffab_pytorch/diffusion.py`` (schedule + three diffusers, injected noise)
* ``ipa``        index-explicit restatement of ``InvariantPointAttentionLayer`` / ``Denoiser``
                 (``diffab_pytorch/diffab_pytorch.py:315-607``) on plain weight dicts
* ``sampler``    the reverse step / ``sample()`` composition (NOT in the reference, SURVEY §3.3;
                 parity for it is "unpinned by the reference" and pinned by our own goldens)
* ``synth``      the synthetic 128-residue patch generator of SURVEY §8(d)
* ``ref_shim``   sys.modules shim that lets the UNMODIFIED reference import in the build
                 container (``/root/reference`` only; it does not exist on the GPU box)

Parity status: every function here except ``sampler`` is pinned against outputs of the reference
itself (``tools/make_goldens.py`` imports the reference through ``ref_shim`` and writes
``tests/golden/*.pt``; ``tests/test_oracle_vs_golden.py`` checks the oracle against them on CPU).
"""
