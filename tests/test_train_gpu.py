"""Training mode on own kernels (BASELINE config 3): SharedMLP with batch-statistics
BatchNorm + max-pool, forward and backward, on the tcgen05 kernels of 3dvlp_b200/mlp_train.py.

Checked (i) against the restatement of its numerics contract (oracle.shared_mlp_pool_mixed: split-bf16
operands, everything else unrounded), (ii) against tests/golden/reference_train.npz -- outputs and gradients of
the REFERENCE's own modules in .train() mode (fp32 torch layers): forward at rtol 1e-3, running statistics at
1e-4, gradients at 2 % relative rms (a max-pool gradient is discontinuous: see _rms_err), (iii) against torch's
own layers on the same device ("torch" backend)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from common import cloud, det_tensor, fill_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_train.npz"))


def _max_norm_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def _rms_err(a, b):
    """Relative root-mean-square error: the measure for gradients.  A max-pool gradient is discontinuous in its
    input -- where two neighbours of a ball are within bf16 round-off of each other the arg-max may legitimately
    differ between two correctly rounded evaluations and one (channel, centre) entry of the gradient moves to
    another slot -- so isolated elements can differ by O(1) while every systematic error shows up here."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-12))


def _rand(name, shape, scale=1.0):
    import zlib
    rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    return torch.from_numpy((rs.normal(size=shape) * scale).astype(np.float32))


def _mlp(pkg, widths, cuda, train=True):
    m = pkg.pytorch_utils.SharedMLP(list(widths), bn=True)
    fill_state_dict(m)
    m = m.to(cuda)
    return m.train() if train else m.eval()


@pytest.mark.parametrize("prec", ["bf16x6", "bf16x3", "bf16"])
@pytest.mark.parametrize("shape,widths,train", [((2, 8, 40, 12), (8, 64, 64, 128), True),
                                                ((3, 131, 33, 16), (131, 128, 128, 256), True),
                                                ((2, 259, 20, 16), (259, 128, 128, 256), True),
                                                ((2, 512, 100, 1), (512, 256, 256), True),
                                                ((2, 4, 300, 64), (4, 64, 64, 128), True),
                                                ((2, 8, 40, 12), (8, 64, 64, 128), False)])
def test_shared_mlp_maxpool_vs_mixed_oracle(pkg, oracle, cuda, shape, widths, train, prec):
    MT = importlib.import_module("3dvlp_b200.mlp_train")
    terms = {"bf16": 1, "bf16x3": 2, "bf16x6": 3}[prec]
    split = terms > 1
    m = _mlp(pkg, widths, cuda, train)
    x = _rand("x%d" % shape[1], shape)
    # repeated rows inside a group (ball-query padding): exact ties of the max-pool must go to the first slot
    if shape[3] >= 4:
        x[:, :, :, -2:] = x[:, :, :, :1]
    dout = _rand("dout", (shape[0], widths[-1], shape[2]))
    blocks = list(m.children())
    ws = [b.conv.weight.detach().cpu() for b in blocks]
    gs = [b.bn.bn.weight.detach().cpu() for b in blocks]
    bs = [b.bn.bn.bias.detach().cpu() for b in blocks]
    running = [(b.bn.bn.running_mean.detach().cpu().clone(), b.bn.bn.running_var.detach().cpu().clone()) for b in blocks]
    want, (dx, dWs, dgs, dbs) = oracle.shared_mlp_pool_mixed(x, ws, gs, bs, running, training=train, dout=dout, terms=terms)

    xg = x.to(cuda).requires_grad_(True)
    assert MT.supported(m, xg)
    with MT.precision(prec):
        got = MT.shared_mlp_maxpool(m, xg)
    assert got.shape == want.shape
    scale = want.abs().max().item()
    # contract tolerance: rtol 1e-3; plain bf16 operands get an absolute floor for the rare one-ulp flips of a
    # hidden activation (2^-9 of one of K terms), the split operands (2^-17) do not need one
    torch.testing.assert_close(got.detach().cpu(), want, rtol=1e-3, atol=(1e-4 if split else 4e-3) * scale)
    (got * dout.to(cuda)).sum().backward()
    tol = 1e-2 if split else 6e-2   # relative rms; arg-max flips between two evaluations of the same contract
    assert _rms_err(xg.grad, dx) < tol, _rms_err(xg.grad, dx)
    if shape[3] >= 4:  # the gradient of a tied group goes to its FIRST slot, none to the repeated rows' pooled share
        assert float(xg.grad[:, :, :, 0].abs().sum()) > float(xg.grad[:, :, :, -1].abs().sum())
    for l, blk in enumerate(blocks):
        assert _rms_err(blk.conv.weight.grad.reshape(dWs[l].shape), dWs[l]) < tol, l
        assert _rms_err(blk.bn.bn.weight.grad, dgs[l]) < tol, l
        assert _rms_err(blk.bn.bn.bias.grad, dbs[l]) < tol, l
        if train:
            torch.testing.assert_close(blk.bn.bn.running_mean.cpu(), running[l][0], rtol=1e-3, atol=1e-4)
            torch.testing.assert_close(blk.bn.bn.running_var.cpu(), running[l][1], rtol=1e-3, atol=1e-4)
            assert int(blk.bn.bn.num_batches_tracked) == 4      # fill_state_dict starts it at 3


def test_shared_mlp_maxpool_vs_torch_layers(pkg, cuda):
    """Same module, same input: the tcgen05 path against torch's own Conv2d / BatchNorm2d / max_pool2d (fp32)."""
    MT = importlib.import_module("3dvlp_b200.mlp_train")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        shape, widths = (2, 131, 64, 32), (131, 128, 128, 256)
        x = _rand("xt", shape).to(cuda)
        dout = _rand("dt", (2, 256, 64)).to(cuda)
        res = {}
        for backend in ("tcgen05", "torch"):
            m = _mlp(pkg, widths, cuda)
            xg = x.clone().requires_grad_(True)
            with MT.mlp_backend(backend):
                if MT.supported(m, xg):
                    assert backend == "tcgen05"
                    y = MT.shared_mlp_maxpool(m, xg)
                else:
                    assert backend == "torch"
                    y = torch.nn.functional.max_pool2d(m(xg), kernel_size=[1, shape[3]]).squeeze(-1)
            (y * dout).sum().backward()
            res[backend] = (y.detach(), xg.grad, [p.grad for p in m.parameters()], [b.clone() for b in m.buffers()])
        a, b = res["tcgen05"], res["torch"]
        assert _max_norm_err(a[0], b[0]) < 1e-3
        assert _rms_err(a[1], b[1]) < 2e-2
        for ga, gb in zip(a[2], b[2]):
            assert _rms_err(ga, gb) < 2e-2
        for ba, bb in zip(a[3], b[3]):
            assert _max_norm_err(ba.float(), bb.float()) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_sa_module_train_step_vs_reference_golden(pkg, gold, cuda):
    from make_golden_train import sa_inputs
    xyz, feats = sa_inputs()
    m = pkg.PointnetSAModuleVotes(npoint=40, radius=0.45, nsample=12, mlp=[5, 64, 64, 128], use_xyz=True, normalize_xyz=True)
    fill_state_dict(m)
    m = m.to(cuda).train()
    x = xyz.to(cuda).requires_grad_(True)
    f = feats.to(cuda).requires_grad_(True)
    nx, nf, ii = m(x, f)
    assert (ii.cpu().numpy() == gold["sa_inds"]).all()
    np.testing.assert_array_equal(nx.detach().cpu().numpy(), gold["sa_new_xyz"])
    np.testing.assert_allclose(nf.detach().cpu().numpy(), gold["sa_features"], rtol=1e-3, atol=1e-4 * float(np.abs(gold["sa_features"]).max()))
    ((nf * det_tensor("w_sa_f", tuple(nf.shape), 1.0).to(cuda)).sum() + (nx * det_tensor("w_sa_x", tuple(nx.shape), 1.0).to(cuda)).sum()).backward()
    assert _rms_err(x.grad, gold["sa_grad_xyz"]) < 2e-2
    assert _rms_err(f.grad, gold["sa_grad_features"]) < 2e-2
    for n, p in m.named_parameters():
        assert _rms_err(p.grad, gold["sa_grad_" + n]) < 2e-2, n
    for n, b in m.named_buffers():
        if n.endswith("num_batches_tracked"):
            assert int(b) == int(gold["sa_buf_" + n])
        else:
            np.testing.assert_allclose(b.cpu().numpy(), gold["sa_buf_" + n], rtol=1e-4, atol=1e-5, err_msg=n)


def test_fp_module_train_step_vs_reference_golden(pkg, gold, cuda):
    from make_golden_train import fp_inputs
    unknown, known, uf, kf = fp_inputs()
    fp = pkg.PointnetFPModule(mlp=[20, 64, 128])
    fill_state_dict(fp)
    fp = fp.to(cuda).train()
    u = uf.to(cuda).requires_grad_(True)
    o = fp(unknown.to(cuda), known.to(cuda), u, kf.to(cuda))
    np.testing.assert_allclose(o.detach().cpu().numpy(), gold["fp_out"], rtol=1e-3, atol=1e-4 * float(np.abs(gold["fp_out"]).max()))
    (o * det_tensor("w_fp", tuple(o.shape), 1.0).to(cuda)).sum().backward()
    assert _rms_err(u.grad, gold["fp_grad_unknow_feats"]) < 2e-2
    for n, p in fp.named_parameters():
        assert _rms_err(p.grad, gold["fp_grad_" + n]) < 2e-2, n
    for n, b in fp.named_buffers():
        if not n.endswith("num_batches_tracked"):
            np.testing.assert_allclose(b.cpu().numpy(), gold["fp_buf_" + n], rtol=1e-4, atol=1e-5, err_msg=n)


def test_backbone_train_step_vs_reference_golden(pkg, gold, cuda):
    """One training step of the whole backbone (train-mode BatchNorm in all 16 layers, backward through FP and SA
    layers, grouping / interpolation scatter-adds) against the reference's own Pointnet2Backbone.

    Sixteen batch-normalised layers deep, this synthetic network amplifies ANY rounding ~7x per SA level, and its
    max-pool / ReLU gradients are discontinuous, so even two fp32 implementations disagree visibly at the end of the
    chain.  The test therefore runs the step twice on the same GPU -- on the tcgen05 kernels and on torch's own
    fp32 layers (TF32 off) -- and holds the tcgen05 path to the same distance from the reference's CPU result as
    torch's fp32 layers reach (x2.5 slack, floors stated below)."""
    from make_golden_train import BB_PARAMS, BB_STATS, backbone_scene
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    try:
        for backend in ("tcgen05", "torch"):
            bb = pkg.Pointnet2Backbone(input_feature_dim=1)
            fill_state_dict(bb)
            bb = bb.to(cuda).train()
            with pkg.mlp_backend(backend):
                d = bb({"point_clouds": backbone_scene().to(cuda)})
                assert (d["fp2_inds"].cpu().numpy() == gold["bb_fp2_inds"]).all()
                fwd = {"sa2": _max_norm_err(d["sa2_features"][:, ::8, ::16].detach(), gold["bb_sa2_features_sub"]),
                       "fp2": _max_norm_err(d["fp2_features"][:, ::8, ::16].detach(), gold["bb_fp2_features_sub"])}
                (d["fp2_features"] * det_tensor("w_bb", tuple(d["fp2_features"].shape), 1.0).to(cuda)).sum().backward()
            named = dict(bb.named_parameters())
            sd = bb.state_dict()
            res[backend] = (fwd, {n: _rms_err(named[n].grad, gold["bb_grad_" + n]) for n in BB_PARAMS},
                            {n: _max_norm_err(sd[n], gold["bb_buf_" + n]) for n in BB_STATS})
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    ours, ref = res["tcgen05"], res["torch"]
    print("backbone train step vs the reference's CPU result -- forward max-norm, gradient relative rms, running stats:")
    print("  tcgen05 kernels :", ours)
    print("  torch fp32 (GPU):", ref)
    assert ours[0]["sa2"] < 1e-3 and ours[0]["fp2"] < max(2.5 * ref[0]["fp2"], 1e-2), ours[0]
    for n in BB_PARAMS:
        assert ours[1][n] < max(2.5 * ref[1][n], 2e-2), (n, ours[1][n], ref[1][n])
    for n in BB_STATS:
        assert ours[2][n] < max(2.5 * ref[2][n], 1e-3), (n, ours[2][n], ref[2][n])


def _split_rows(a, terms):
    """fp32 (rows, c) -> bf16 rows in the `terms` layout of the training kernels: [h], [h|m], [h|m|l]."""
    h = a.bfloat16()
    r1 = a - h.float()
    m = r1.bfloat16()
    lo = (r1 - m.float()).bfloat16()
    t = (h, m, lo)
    order = tuple(range(terms))
    return torch.cat([t[i] for i in order], dim=1).contiguous(), t


@pytest.fixture(params=[1, 0], ids=["tma", "cpasync"])
def wgrad_variant(request, pkg):
    """Both operand-copy variants of the weight-gradient kernel (option "wgrad_tma"): tiled TMA through tensor
    maps (the default) and the cp.async fallback; the launch statistics prove which one ran."""
    name = "wgrad_tma_launches" if request.param else "wgrad_cpasync_launches"
    pkg.set_option("wgrad_tma", request.param)
    before = pkg.get_stat(name)
    yield request.param
    pkg.set_option("wgrad_tma", 1)
    assert pkg.get_stat(name) > before, "%s did not move: the other kernel variant ran" % name


@pytest.mark.parametrize("terms", [1, 2, 3])
@pytest.mark.parametrize("rows,c_out,c_in", [(5000, 64, 4), (4100, 128, 64), (3000, 256, 259), (130, 64, 131)])
def test_wgrad_split_one_launch(pkg, cuda, wgrad_variant, terms, rows, c_out, c_in):
    """pn2_wgrad_split_bf16: every leading (dy term, x term) product of the split-precision weight gradient in one
    launch, accumulated in one TMEM tile -- against the fp64 sum of the same bf16 term products (the contraction is
    exact up to fp32 accumulation) and against one pn2_wgrad_bf16 call per product."""
    import ctypes
    L = pkg._lib
    g = torch.Generator().manual_seed(rows + c_in)
    dy32 = (torch.randn(rows, c_out, generator=g) * 0.3).to(cuda)
    kx = (c_in + 15) // 16 * 16
    x32 = torch.zeros(rows, kx, device=cuda)
    x32[:, :c_in] = torch.rand(rows, c_in, generator=g).to(cuda) * 2 - 0.5
    dy, dyt = _split_rows(dy32, terms)
    x, xt = _split_rows(x32, terms)
    prods = {1: [(0, 0)], 2: [(0, 0), (1, 0), (0, 1)], 3: [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (1, 1)]}[terms]
    want = sum(dyt[i].double().t() @ xt[j].double()[:, :c_in] for i, j in prods)
    nblk = terms
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    dw = torch.zeros(c_out, c_in, device=cuda)
    L.call("pn2_wgrad_split_bf16", ctypes.c_longlong(rows), c_out, c_in, nblk * c_out, nblk * kx, L.ptr(dy), L.ptr(x),
           terms, L.ptr(dw), st, device=cuda)
    scale = float(want.abs().max())
    assert float((dw.double() - want).abs().max()) <= 2e-5 * scale
    first = {0: 0, 1: 1, 2: 2}
    dw6 = torch.zeros(c_out, c_in, device=cuda)
    for i, j in prods:
        L.call("pn2_wgrad_bf16", ctypes.c_longlong(rows), c_out, c_in, nblk * c_out, nblk * kx,
               ctypes.c_void_p(dy.data_ptr() + 2 * first[i] * c_out), ctypes.c_void_p(x.data_ptr() + 2 * first[j] * kx),
               L.ptr(dw6), st, device=cuda)
    assert float((dw - dw6).abs().max()) <= 2e-5 * scale


def test_graphed_train_step_matches_eager(pkg, cuda):
    """graph.GraphedTrainStep: the whole training iteration of the detector (train-mode forward on the tcgen05
    training path, detection + DIoU losses, backward, Adam) captured as ONE CUDA graph.  From the same model /
    optimizer state a replay and an eager iteration must produce the same loss, gradients, running statistics
    and updated parameters (up to the order of the fp32 atomics of the scatter-add / weight gradients)."""
    import types
    syn = importlib.import_module("3dvlp_b200.synthetic")
    LD = importlib.import_module("3dvlp_b200.loss_detection")
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    msa = syn.mean_size_arr()
    cfg = types.SimpleNamespace(num_class=18, num_heading_bin=1, num_size_cluster=18, mean_size_arr=msa,
                                mean_size_arr_tensor=torch.from_numpy(msa).to(cuda))
    host = syn.make_batch(3, 2, 8000, 1)
    pc = host.to(cuda)
    labels = {k: v.to(cuda) for k, v in syn.make_labels(host, seed=7, lang_num_max=4).items()}
    cluster_ref = torch.randn(2 * 4, 256, generator=torch.Generator().manual_seed(5)).to(cuda)
    torch.manual_seed(11)
    det = pkg.VoteNetDetector(input_feature_dim=1, mean_size_arr=msa).to(cuda).train()
    opt = torch.optim.Adam(det.parameters(), lr=1e-3, capturable=True)

    def loss_fn():
        d = det({"point_clouds": pc})
        d.update(labels)
        d.update({"epoch": 10, "istrain": 1, "random": 0.7, "cluster_ref": cluster_ref})
        d = LD.get_detection_loss(d, cfg)
        d, ref_loss, _, _ = LG.compute_diou_loss(d, cfg)
        return d["vote_loss"] + 0.1 * d["objectness_loss"] + d["box_loss"] + 0.3 * ref_loss + d["diou_loss"]

    with pytest.raises(RuntimeError, match="capturable"):
        pkg.GraphedTrainStep(loss_fn, torch.optim.Adam(det.parameters(), lr=1e-3))
    g = pkg.GraphedTrainStep(loss_fn, opt, warmup=2)
    params = dict(det.named_parameters())
    state = [t for st in opt.state.values() for t in st.values() if torch.is_tensor(t)]
    tensors = list(params.values()) + list(det.buffers()) + state
    snap = [t.detach().clone() for t in tensors]

    loss_g = g.step().clone()
    grads_g = {n: p.grad.detach().clone() for n, p in params.items() if p.grad is not None}
    after_g = [t.detach().clone() for t in tensors]
    with torch.no_grad():
        for t, s in zip(tensors, snap):
            t.copy_(s)
    loss = loss_fn()
    names = [n for n in params if n in grads_g]
    grads_e = dict(zip(names, torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)))
    loss_e = float(loss.detach())
    assert abs(loss_e - float(loss_g)) <= 1e-4 * abs(loss_e), (loss_e, float(loss_g))
    assert len(grads_g) > 60  # backbone, voting and proposal parameters received a gradient inside the graph
    worst = 0.0
    for n in names:
        if grads_e[n] is None:
            continue
        worst = max(worst, _rms_err(grads_g[n], grads_e[n].cpu().numpy()))
    assert worst < 1e-3, worst
    # the eager optimizer step from the restored state lands on the replay's parameters / statistics
    for n in names:
        params[n].grad = grads_e[n]
    opt.step()
    n_par = len(params)
    for i, (t, a) in enumerate(zip(tensors, after_g)):
        if i < n_par:  # Adam turns gradient noise of near-zero entries into +-lr steps: compare in units of lr
            assert float((t.detach() - a).abs().max()) <= 2.5e-3, i
        elif t.dtype.is_floating_point and i < n_par + len(list(det.buffers())):
            np.testing.assert_allclose(t.cpu().numpy(), a.cpu().numpy(), rtol=1e-4, atol=1e-6)
    assert pkg.tensorcore.poll_device_error(cuda) == 0
