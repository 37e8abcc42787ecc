import sys, json, torch
sys.path.insert(0, ".")
from bench_extra import cache_to_features_block
print(json.dumps(cache_to_features_block(torch.device("cuda:0")), indent=1))
