"""One fit+predict pass at size N (default 40000) -- the target process of the ncu captures in tools/runs/profile_1gpu.sh."""
import sys

import torch

sys.path.insert(0, ".")
from battgp_b200 import engine as E
from battgp_b200.synth import query_grid, synth_field_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
x, y = synth_field_data(n, 0)
xd, yd, xq = torch.tensor(x, device=dev), torch.tensor(y, device=dev), torch.tensor(query_grid(x), device=dev)
K = E.alloc_matrix(n + xq.shape[0], n, dev)
for _ in range(reps):
    st = E.fit(E.battgp_spec(), xd, yd, 2.33e-6, K_out=K, xq=xq)
    mean, var = E.predict(st, xq)
torch.cuda.synchronize()
print("one_fit", n, st.lml, float(mean[0]), float(var[0]), E.get_engine(dev).launches)
