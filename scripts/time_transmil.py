import json, sys
sys.path.insert(0, "/root/repo")
import torch
from bench_extra import aggregators_block
print(json.dumps(aggregators_block(torch.device("cuda:0"), 6543.0)["transmil"]))
