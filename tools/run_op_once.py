"""Run one standalone op at a large shape a few times: the target command for `ncu --set full`
captures of the bandwidth-bound kernels (see profiles/README.md).
Usage: python tools/run_op_once.py group_points|group_points_grad|three_interpolate|three_interpolate_grad|three_nn"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
ext = importlib.import_module("3dvlp_b200._ext")
syn = importlib.import_module("3dvlp_b200.synthetic")
op = sys.argv[1] if len(sys.argv) > 1 else "group_points"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if op.startswith("group_points"):
    B, C, N, M, S = 8, 128, 40000, 2048, 64
    xyz = syn.make_batch(7, B, N, 1)[..., :3].contiguous().to(dev)
    ind = ext.furthest_point_sampling(xyz, M)
    nx = torch.gather(xyz, 1, ind.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx = ext.ball_query(nx, xyz, 0.2, S)
    feats = torch.rand(B, C, N, generator=g).to(dev)
    go = torch.rand(B, C, M, S, generator=g).to(dev)
    for _ in range(3):
        out = ext.group_points(feats, idx) if op == "group_points" else ext.group_points_grad(go, idx, N)
elif op.startswith("three_interpolate"):
    B, C, m, n = 32, 256, 2048, 40000
    known = torch.rand(B, C, m, generator=g).to(dev)
    d2, idx = ext.three_nn(torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, m, 3, generator=g).to(dev))
    w = 1.0 / (torch.sqrt(d2) + 1e-8)
    w = (w / w.sum(2, keepdim=True)).contiguous()
    go = torch.rand(B, C, n, generator=g).to(dev)
    for _ in range(3):
        out = ext.three_interpolate(known, idx, w) if op == "three_interpolate" else ext.three_interpolate_grad(go, idx, w, m)
elif op == "three_nn":
    B, n, m = 8, 20000, 2048
    u = syn.make_batch(3, B, n, 1)[..., :3].contiguous().to(dev)
    k = u[:, :m].contiguous()
    for _ in range(3):
        out = ext.three_nn(u, k)
torch.cuda.synchronize()
print("ok", op)
