"""bench.py's int8 config-4 sub-dict alone (decode bs1 / bs8 linears in one graph, e2e through the decoder, prefill):
for tuning the int8 decode kernel without the rest of the bench."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
out = bench.int8_config4(torch, dev, bench.load_peaks())
print(json.dumps({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "how"}) for k, v in out.items()}))
