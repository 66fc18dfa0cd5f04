"""Development probe: BASELINE.json configs[2] -- Matern-5/2-ARD(I,SOC,T) + Periodic(t), N points, a few Adam iterations of
the LML through the GPyTorch-shaped facade (mll forward + analytic backward).  Prints one JSON line."""
import json, sys, time
import torch
sys.path.insert(0, ".")
import battgp_b200.shim as shim
shim.install(force=True)
import gpytorch
from battgp_b200.synth import synth_field_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
x, y = synth_field_data(n, 0)
xt, yt = torch.tensor(x, device=dev), torch.tensor(y, device=dev)


class M(gpytorch.models.ExactGP):
    def __init__(self, tx, ty):
        super().__init__(tx, ty, gpytorch.likelihoods.GaussianLikelihood(noise_constraint=gpytorch.constraints.Interval(0.0, 1e5)))
        self.mean_module = gpytorch.means.ZeroMean()
        km = gpytorch.kernels.MaternKernel(nu=2.5, ard_num_dims=3, active_dims=[1, 2, 3])
        kp = gpytorch.kernels.PeriodicKernel(active_dims=[0])
        self.covar_module = gpytorch.kernels.ScaleKernel(km) + gpytorch.kernels.ScaleKernel(kp)
        self.to(tx.device)

    def forward(self, xx):
        return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))


m = M(xt, yt)
m.likelihood.noise = torch.tensor([2.33e-6], device=dev)
m.covar_module.kernels[0].outputscale = torch.tensor(0.0099, device=dev)
m.covar_module.kernels[0].base_kernel.lengthscale = torch.tensor([12.11, 33.75, 45.14], device=dev)
m.covar_module.kernels[1].outputscale = torch.tensor(1e-4, device=dev)
m.covar_module.kernels[1].base_kernel.lengthscale = torch.tensor(1.0, device=dev)
m.covar_module.kernels[1].base_kernel.period_length = torch.tensor(1.0, device=dev)
m.train(); m.likelihood.train()
mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)
opt = torch.optim.Adam(m.parameters(), lr=0.1)
times, losses = [], []
for it in range(iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    loss = -mll(m(xt), yt)
    loss.backward()
    opt.step()
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    losses.append(float(loss) * n)
print(json.dumps({"op": "train_iter", "kernel": "matern52_ard+periodic", "n": n, "iters": iters, "seconds_per_iter": times,
                  "loss_times_n": losses, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
