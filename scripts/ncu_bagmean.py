import sys
sys.path.insert(0, "/root/repo")
import torch
from stamp_b200.mlp import bag_mean
x = torch.randn(64, 4096, 1024, device="cuda").half()
for _ in range(3):
    bag_mean(x)
torch.cuda.synchronize()
