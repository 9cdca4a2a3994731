import os, sys, json, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
shapes = torch.load("/root/repo/tests/golden/state_shapes.pt", weights_only=False)
r = bench.measure_train_step(dev, None, 1, shapes, steps=10)
print("train step", r["ms_per_step"], "ms", r["loss_finite"])
