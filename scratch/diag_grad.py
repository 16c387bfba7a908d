import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'links-3d-human-pose-estimation_b200')
from links_b200.mlp import MlpSet
from oracle import nets as ON, nets_bf16 as ONB, steps as OS


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


M = 200; nj = (7, 10)
params = [ON.init_lifter_params(7, 11), ON.init_lifter_params(10, 12)]
mlp = MlpSet("lifter", [14, 20], [{"downscale": 7, "angles": 1}, {"downscale": 10, "angles": 1}], M, n_passes=1)
mlp.load_state_dicts(params)
g = torch.Generator().manual_seed(0)
xs = [torch.randn(M, 2 * n, generator=g) * 0.15 for n in nj]
st = torch.cuda.current_stream().cuda_stream
for s in range(2):
    idx = torch.arange(2 * nj[s], dtype=torch.int32, device="cuda"); xd = xs[s].cuda()
    mlp.lib.links_pack_rows(xd.data_ptr(), xd.stride(0), M, idx.data_ptr(), 2 * nj[s], 1, mlp.x0[0][s].data_ptr(), mlp.x0T[s].data_ptr(), mlp.ldT, 0, st)
mlp.run(mlp.forward_plan(0))
ups = []
g2 = torch.Generator().manual_seed(1)
for s in range(2):
    gd = (torch.randn(M, nj[s], generator=g2) * 0.1).bfloat16(); ga = (torch.randn(M, 1, generator=g2) * 0.1).bfloat16()
    ups.append((gd, ga)); G, GT = mlp.G[0][s], mlp.GT[s]
    G["downscale"].zero_(); G["angles"].zero_()
    G["downscale"][:, :nj[s]] = gd.cuda(); G["angles"][:, :1] = ga.cuda()
    GT["downscale"][:, :M] = gd.cuda().t(); GT["angles"][:, :M] = ga.cuda().t()
mlp.run(mlp.backward_plan(0, need_input_grad=True)); mlp.run(mlp.wgrad_plan()); torch.cuda.synchronize()
s = 0
# twin with hooks to capture intermediate activations / grads
p = OS.params_require_grad(params[s]); x = xs[s].clone().requires_grad_(True)
xd, xa = ONB.lifter_forward(x, p); ((xd * ups[s][0].float()).sum() + (xa * ups[s][1].float()).sum()).backward()
print("gW vs twin", {n: round(rel(mlp.nets[s].layers[n].gW.cpu(), p[n + ".weight"].grad), 4) for n in mlp.layer_names})
print("gb vs twin", {n: round(rel(mlp.nets[s].layers[n].gb.cpu(), p[n + ".bias"].grad), 4) for n in mlp.layer_names})
print("din", rel(mlp.din[0][s][:, :14].cpu(), x.grad))
# forward activations vs twin: recompute twin activations explicitly
import oracle.nets as N32
rb = ONB.rb
W = {k: rb(v.detach()) for k, v in params[s].items()}
xin = rb(xs[s])
h0 = rb(xin @ W["upscale.weight"].t() + params[s]["upscale.bias"])
print("h0", rel(mlp.act[0][s]["h0"].float().cpu(), h0), "exact frac", (mlp.act[0][s]["h0"].float().cpu() == h0).float().mean().item())
a1 = rb(N32.leaky(h0 @ W["res_common.l1.weight"].t() + params[s]["res_common.l1.bias"]))
print("common.a1", rel(mlp.act[0][s]["res_common.a1"].float().cpu(), a1), (mlp.act[0][s]["res_common.a1"].float().cpu() == a1).float().mean().item())
z2 = a1 @ W["res_common.l2.weight"].t() + params[s]["res_common.l2.bias"]
y = rb(N32.leaky(N32.leaky(z2) + h0))
print("common.y", rel(mlp.act[0][s]["res_common.y"].float().cpu(), y), (mlp.act[0][s]["res_common.y"].float().cpu() == y).float().mean().item())
sg = mlp.sign[0][s]["res_common"].cpu()
bits = ((sg.unsqueeze(-1) >> torch.arange(32, dtype=torch.int32)) & 1).reshape(M, 1024).bool()
print("sign mismatch frac", (bits != ~(z2 > 0)).float().mean().item())
# backward intermediate: G of last block
print("G pose3.l2 nonzero frac", (mlp.G[0][s]["res_pose3.l2"].float() != 0).float().mean().item())
