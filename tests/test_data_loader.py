"""Host logic of the preprocessed-patch reader: the file layout of the reference's
preprocess_pdb.py:67-80, padding of ragged patches, error behaviour.  CPU only."""
import pytest
import torch

from diffab_pytorch_b200 import data, synth


def _write_patch(path, L, seed, with_gen=False, drop=None):
    b = synth.make_patches(1, L, seed=seed, with_distmat=False, cdr=(5, 9))
    d = {k: b[k] for k in data.PATCH_KEYS if k in b}
    d["backbone_dihedrals_mask"] = torch.ones(1, L, 3, dtype=torch.bool)
    if with_gen:
        d["generation_mask"] = b["generation_mask"]
    if drop:
        d.pop(drop)
    torch.save(d, path)
    return b


def test_ragged_patches_are_padded_with_masked_residues(tmp_path):
    lengths = [24, 32, 17]
    paths, raw = [], []
    for i, L in enumerate(lengths):
        p = str(tmp_path / f"patch{i}.pt")
        raw.append(_write_patch(p, L, seed=i))
        paths.append(p)
    ds = data.PatchDataset(paths, generation_mask_fn=data.span_mask([(5, 9)]))
    assert len(ds) == 3
    batch = data.collate_patches([ds[i] for i in range(3)])
    assert batch["xyz"].shape == (3, 32, 15, 3) and batch["pairwise_dihedrals"].shape == (3, 32, 32, 2)
    assert batch["distmat"].shape == (3, 32, 32, 15, 15)
    for i, L in enumerate(lengths):
        for k in ("xyz", "orientations", "seq_idx", "chain_idx", "backbone_dihedrals"):
            assert torch.equal(batch[k][i, :L], raw[i][k][0]), k
        assert torch.equal(batch["pairwise_dihedrals"][i, :L, :L], raw[i]["pairwise_dihedrals"][0])
        assert bool(batch["residue_mask"][i, :L].all()) and not bool(batch["residue_mask"][i, L:].any())
        assert not bool(batch["atom_mask"][i, L:].any()) and not bool(batch["generation_mask"][i, L:].any())
        assert bool((batch["seq_idx"][i, L:] == data.AA_UNK).all()) and bool((batch["chain_idx"][i, L:] == 0).all())
        assert torch.equal(batch["generation_mask"][i, :L], raw[i]["generation_mask"][0])
        # residue numbering stays strictly increasing through the padding
        assert bool((batch["residue_idx"][i, 1:] > batch["residue_idx"][i, :-1]).all())
        # distances of real atoms are those of the coordinates
        ref = synth.pairwise_atom_distances(raw[i]["xyz"])[0]
        assert torch.allclose(batch["distmat"][i, :L, :L], ref)
    fixed = data.collate_patches([ds[0], ds[2]], length=40, with_distmat=False)
    assert fixed["xyz"].shape == (2, 40, 15, 3) and "distmat" not in fixed


def test_stored_generation_mask_and_errors(tmp_path):
    p = str(tmp_path / "a.pt")
    b = _write_patch(p, 16, seed=3, with_gen=True)
    item = data.PatchDataset([p])[0]
    assert torch.equal(item["generation_mask"], b["generation_mask"][0])
    q = str(tmp_path / "b.pt")
    _write_patch(q, 16, seed=4)
    with pytest.raises(ValueError, match="generation_mask"):
        data.PatchDataset([q])[0]
    r = str(tmp_path / "c.pt")
    _write_patch(r, 16, seed=5, drop="orientations")
    with pytest.raises(ValueError, match="missing keys"):
        data.load_patch(r)
    with pytest.raises(ValueError, match="shorter"):
        data.collate_patches([item], length=8)
