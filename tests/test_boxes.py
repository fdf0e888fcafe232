"""Box decode, IoU / DIoU, the DIoU reference loss and the OCC/OSC contrastive module
(SURVEY.md 8f ranks 3-4): oracle vs the reference's golden vectors on CPU, CUDA path vs oracle /
goldens on the GPU, and the all-gather variant of the contrastive loss on two gloo ranks."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_boxes import box_cases  # noqa: E402  (deterministic inputs; no reference import at module level)
from common import det_tensor  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_boxes.npz"))


@pytest.fixture(scope="module")
def OB():
    from oracle import boxes
    return boxes


MEAN_SIZES = det_tensor("mean_sizes", (18, 3), 0.6).abs() + 0.3


def grounding_inputs(seed, B=3, K=32, Lmax=5, epoch=10, istrain=1, rnd=0.3):
    """A data_dict shaped like the one compute_diou_loss / ContrastModule read."""
    g = torch.Generator().manual_seed(seed)
    gt_center = torch.rand(B, Lmax, 3, generator=g) * 3
    size_cls = torch.randint(0, 18, (B, Lmax), generator=g)
    size_res = (torch.rand(B, Lmax, 3, generator=g) - 0.5) * 0.1
    gt_size = MEAN_SIZES[size_cls] + size_res
    pred_center = torch.rand(B, K, 3, generator=g) * 3
    pred_size = torch.rand(B, K, 3, generator=g) * 0.8 + 0.2
    # plant proposals on the targets: several good ones for sentence 0, one for sentence 1, none further
    for b in range(B):
        for j, n_good in ((0, 4), (1, 1), (2, 2)):
            for t in range(n_good):
                k = 5 * j + t
                pred_center[b, k] = gt_center[b, j] + (torch.rand(3, generator=g) - 0.5) * 0.08
                pred_size[b, k] = gt_size[b, j] * (1 + (torch.rand(3, generator=g) - 0.5) * 0.1)
    d = {
        "pred_center": pred_center, "pred_size": pred_size,
        "pred_center_reg": (torch.rand(B, Lmax, K, 3, generator=g) - 0.5) * 0.04,
        "pred_size_reg": (torch.rand(B, Lmax, K, 3, generator=g) - 0.5) * 0.04,
        "ref_center_label_list": gt_center, "ref_size_class_label_list": size_cls,
        "ref_size_residual_label_list": size_res,
        "ref_heading_class_label_list": torch.zeros(B, Lmax, dtype=torch.long),
        "ref_heading_residual_label_list": torch.zeros(B, Lmax),
        "aggregated_vote_features": torch.zeros(B, K, 128),
        "lang_num": torch.tensor([Lmax, 3, 1][:B]),
        "cluster_ref": torch.randn(B * Lmax, K, generator=g),
        "objectness_scores": torch.randn(B, K, 2, generator=g),
        "istrain": torch.tensor([istrain]), "random": rnd, "epoch": epoch,
        "bbox_feature": torch.randn(B, K, 128, generator=g),
        "lang_emb": torch.randn(B * Lmax, 128, generator=g),
    }
    d["objectness_scores"][:, :12, 1] += 3.0  # the planted proposals count as objects
    return d


def to_dev(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


CONFIG = types.SimpleNamespace(mean_size_arr_tensor=MEAN_SIZES)


# ------------------------------------------------------------------------------------------ CPU
def test_oracle_diou_matches_reference_golden(OB, gold):
    c1, s1, c2, s2 = [t.clone().requires_grad_(True) for t in box_cases()]
    iou, diou = OB.box3d_diou_batch_tensor(c1, s1, c2, s2)
    assert np.array_equal(iou.detach().numpy(), gold["diou_iou"], equal_nan=True)
    assert np.array_equal(diou.detach().numpy(), gold["diou_diou"], equal_nan=True)
    (diou * det_tensor("gd", diou.shape, 1.0)).sum().backward()
    for n, t in zip(("c1", "s1", "c2", "s2"), (c1, s1, c2, s2)):
        np.testing.assert_allclose(t.grad.numpy(), gold["diou_gdiou_" + n], rtol=1e-6, atol=1e-7)
    x, t = det_tensor("rank_x", (5, 16), 3.0), det_tensor("rank_t", (5, 16), 0.5).abs()
    np.testing.assert_allclose(OB.softmax_ranking_loss(x, t).numpy(), gold["softmax_ranking"], rtol=1e-6)


def _contrast_terms(pkg_contrast, OB, d):
    """Embeddings + IoU targets for occ_osc_losses from oracle pieces (CPU)."""
    import torch.nn.functional as F
    B, K = d["pred_center"].shape[:2]
    Lmax = d["ref_center_label_list"].shape[1]
    gt_size = MEAN_SIZES[d["ref_size_class_label_list"]] + d["ref_size_residual_label_list"] + 1e-2
    c1 = d["pred_center"][:, None].expand(B, Lmax, K, 3).reshape(-1, 3)
    s1 = d["pred_size"][:, None].expand(B, Lmax, K, 3).reshape(-1, 3)
    c2 = d["ref_center_label_list"][:, :, None].expand(B, Lmax, K, 3).reshape(-1, 3)
    s2 = gt_size[:, :, None].expand(B, Lmax, K, 3).reshape(-1, 3)
    ious, _ = OB.box3d_diou_batch_tensor(c1, s1, c2, s2)
    obj = d["objectness_scores"].max(2)[1].bool()
    target = ((ious.view(B, Lmax, K) > 0.25) & obj[:, None]).float()
    valid = torch.arange(Lmax)[None] < d["lang_num"][:, None]
    g = torch.Generator().manual_seed(5)
    w = [torch.randn(128, 128, generator=g) * 0.1 for _ in range(3)]
    text = F.normalize(d["lang_emb"].view(B, Lmax, -1) @ w[1].t(), dim=-1)
    box = F.normalize(d["bbox_feature"] @ w[0].t(), dim=-1)
    box_iou = F.normalize(d["bbox_feature"] @ w[2].t(), dim=-1)
    return w, text, box, box_iou, target, obj, valid


def test_contrast_batched_equals_reference_loops_cpu(pkg, OB):
    C = importlib.import_module("3dvlp_b200.contrast")
    d = grounding_inputs(3, epoch=60)
    w, text, box, box_iou, target, obj, valid = _contrast_terms(C, OB, d)
    lang, iou = C.occ_osc_losses(text, box, box_iou, target, obj, valid)
    want_lang, want_iou = OB.contrast_loop(d, MEAN_SIZES, w[0], w[1], w[2])
    torch.testing.assert_close(lang, want_lang, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(iou, want_iou, rtol=1e-5, atol=1e-6)
    assert float(want_lang) > 0 and float(want_iou) > 0
    # building blocks keep the reference's definitions
    x, t = torch.randn(4, 6), torch.rand(4, 6)
    torch.testing.assert_close(C.SoftCrossEntropy(x, t), OB._soft_ce(x, t))
    torch.testing.assert_close(C.NCELoss()(x[:, :4], t[:, :4]), OB._nce(x[:, :4], t[:, :4]))
    assert sorted(C.ContrastModule(None).state_dict()) == ["nce_loss.tau", "pc_proj.weight", "pc_proj_iou.0.weight",
                                                            "text_proj.weight"]
    cb = C.create_box_batch(torch.tensor([[1.0, 2, 3]]), torch.tensor([[2.0, 4, 6]]))
    assert cb.shape == (1, 8, 3) and torch.equal(cb[0, 0], torch.tensor([0.0, 0, 0])) and torch.equal(cb[0, 6], torch.tensor([2.0, 4, 6]))


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import boxes as OB
        C = importlib.import_module("3dvlp_b200.contrast")
        parts = [_contrast_terms(C, OB, grounding_inputs(40 + r, B=2, epoch=60)) for r in range(world)]
        _, text, box, box_iou, target, obj, valid = parts[rank]
        box = box.clone().requires_grad_(True)
        lang, iou = C.occ_osc_losses(text, box, box_iou, target, obj, valid, gather=True)
        # expected: this rank's sentences against the kept objects of ALL ranks' scenes
        all_box = torch.cat([p[2] for p in parts]).reshape(-1, 128)
        all_obj = torch.cat([p[5] for p in parts]).reshape(-1)
        B, K = obj.shape
        want = 0.
        for b in range(B):
            for j in range(text.shape[1]):
                if valid[b, j]:
                    sim = text[b, j] @ all_box[all_obj].t()
                    logp = torch.log_softmax(sim, -1)
                    tgt = torch.zeros(all_obj.numel())
                    tgt[(rank * B + b) * K:(rank * B + b + 1) * K] = target[b, j]
                    want = want + (-(logp * tgt[all_obj]).sum() / obj[b].sum() / 2)
        torch.testing.assert_close(lang, want / B, rtol=1e-5, atol=1e-6)
        local_lang, local_iou = C.occ_osc_losses(text, box, box_iou, target, obj, valid, gather=False)
        torch.testing.assert_close(iou, local_iou)          # the object-object term stays per scene
        assert float(lang) > float(local_lang)               # more negatives -> larger loss
        lang.backward()
        assert box.grad is not None and float(box.grad.abs().sum()) > 0
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def test_contrast_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_diou_kernel_bit_exact_vs_reference_golden(pkg, cuda, gold):
    BU = importlib.import_module("3dvlp_b200.box_util")
    c1, s1, c2, s2 = [t.to(cuda).requires_grad_(True) for t in box_cases()]
    iou, diou = BU.box3d_diou_batch_tensor(c1, s1, c2, s2)
    assert np.array_equal(iou.detach().cpu().numpy(), gold["diou_iou"], equal_nan=True)
    assert np.array_equal(diou.detach().cpu().numpy(), gold["diou_diou"], equal_nan=True)
    gi, gd = det_tensor("gi", iou.shape, 1.0).to(cuda), det_tensor("gd", diou.shape, 1.0).to(cuda)
    (iou * gi).sum().backward(retain_graph=True)
    for n, t in zip(("c1", "s1", "c2", "s2"), (c1, s1, c2, s2)):
        np.testing.assert_allclose(t.grad.cpu().numpy(), gold["diou_giou_" + n], rtol=2e-4, atol=2e-5, err_msg=n)
        t.grad = None
    (diou * gd).sum().backward()
    for n, t in zip(("c1", "s1", "c2", "s2"), (c1, s1, c2, s2)):
        np.testing.assert_allclose(t.grad.cpu().numpy(), gold["diou_gdiou_" + n], rtol=2e-4, atol=2e-5, err_msg=n)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        BU.box3d_diou_batch_tensor(*box_cases())


@pytest.mark.gpu
def test_diou_pairs_and_grad(pkg, cuda, OB):
    BU = importlib.import_module("3dvlp_b200.box_util")
    d = grounding_inputs(7, B=4, K=64, Lmax=6)
    gt_size = MEAN_SIZES[d["ref_size_class_label_list"]] + d["ref_size_residual_label_list"]
    pc = d["pred_center"].to(cuda).requires_grad_(True)
    ps = d["pred_size"].to(cuda).requires_grad_(True)
    iou, diou = BU.box3d_diou_pairs(pc, ps, d["ref_center_label_list"].to(cuda), gt_size.to(cuda))
    B, K, Lmax = 4, 64, 6
    rc = d["pred_center"].clone().requires_grad_(True)
    rs = d["pred_size"].clone().requires_grad_(True)
    wi, wd = OB.box3d_diou_batch_tensor(rc[:, None].expand(B, Lmax, K, 3).reshape(-1, 3),
                                        rs[:, None].expand(B, Lmax, K, 3).reshape(-1, 3),
                                        d["ref_center_label_list"][:, :, None].expand(B, Lmax, K, 3).reshape(-1, 3),
                                        gt_size[:, :, None].expand(B, Lmax, K, 3).reshape(-1, 3))
    assert torch.equal(iou.cpu(), wi.view(B, Lmax, K)) and torch.equal(diou.cpu(), wd.view(B, Lmax, K))
    assert float((iou > 0.25).sum()) >= 10
    gi, gd = torch.randn(B, Lmax, K), torch.randn(B, Lmax, K)
    ((iou * gi.to(cuda)).sum() + (diou * gd.to(cuda)).sum()).backward()
    ((wi.view(B, Lmax, K) * gi).sum() + (wd.view(B, Lmax, K) * gd).sum()).backward()
    torch.testing.assert_close(pc.grad.cpu(), rc.grad, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(ps.grad.cpu(), rs.grad, rtol=2e-4, atol=2e-5)


@pytest.mark.gpu
def test_box_decode_on_device(pkg, cuda, gold):
    BU = importlib.import_module("3dvlp_b200.box_util")
    size = (det_tensor("bs", (2, 7, 3), 1.0).abs() + 0.2).to(cuda)
    heading, center = det_tensor("bh", (2, 7), 3.0).to(cuda), det_tensor("bc", (2, 7, 3), 2.0).to(cuda)
    np.testing.assert_allclose(BU.get_3d_box_batch(size, heading, center).cpu().numpy(), gold["corners"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(BU.rotz_batch_pytorch(heading).cpu().numpy(), gold["rotz"], rtol=1e-6, atol=1e-6)
    B, K, NH = 2, 9, 4
    d = {"aggregated_vote_xyz": det_tensor("dx", (B, K, 3), 2.0).to(cuda),
         "heading_scores": det_tensor("dhs", (B, K, NH), 1.0).to(cuda),
         "heading_residuals": det_tensor("dhr", (B, K, NH), 0.3).to(cuda),
         "rois": (det_tensor("droi", (B, K, 6), 0.7).abs() + 0.05).to(cuda)}
    BU.decode_pred_box(d, NH)
    for key, g in (("pred_heading", "dec_heading"), ("pred_size", "dec_size"), ("pred_center", "dec_center"),
                   ("pred_bbox_corner", "dec_corners")):
        assert d[key].is_cuda
        np.testing.assert_allclose(d[key].cpu().numpy(), gold[g], rtol=1e-5, atol=1e-5, err_msg=key)


@pytest.mark.gpu
@pytest.mark.parametrize("epoch,istrain,rnd,reg", [(10, 1, 0.3, False), (10, 1, 0.9, False), (60, 0, 0.3, False),
                                                   (10, 1, 0.3, True), (60, 1, 0.2, True)])
def test_compute_diou_loss_equals_reference_loops(pkg, cuda, OB, epoch, istrain, rnd, reg):
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    d = grounding_inputs(11 + epoch, epoch=epoch, istrain=istrain, rnd=rnd)
    want_loss, want_diou, want_labels, r25, r5 = OB.compute_diou_loss_loop(d, MEAN_SIZES, use_reg_head=reg)
    dd = to_dev(d, cuda)
    dd["pred_center"].requires_grad_(True)
    out, loss, preds, labels = LG.compute_diou_loss(dd, CONFIG, use_reg_head=reg)
    assert torch.equal(labels.cpu(), want_labels)
    assert float(want_labels.sum()) >= 3
    torch.testing.assert_close(loss.cpu(), want_loss, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["diou_loss"].cpu(), want_diou, rtol=1e-5, atol=1e-6)
    assert abs(float(out["max_iou_rate_0.25"]) - r25) < 1e-6 and abs(float(out["max_iou_rate_0.5"]) - r5) < 1e-6
    assert preds.shape == (3, 5, 32)
    out["diou_loss"].backward()
    assert float(dd["pred_center"].grad.abs().sum()) > 0


@pytest.mark.gpu
def test_contrast_module_equals_reference_loops(pkg, cuda, OB):
    C = importlib.import_module("3dvlp_b200.contrast")
    d = grounding_inputs(21, epoch=60)
    m = C.ContrastModule(CONFIG).to(cuda)
    want_lang, want_iou = OB.contrast_loop(d, MEAN_SIZES, m.pc_proj.weight.detach().cpu(), m.text_proj.weight.detach().cpu(),
                                           m.pc_proj_iou[0].weight.detach().cpu())
    out = m(to_dev(d, cuda))
    torch.testing.assert_close(out["lang_con_loss"].cpu(), want_lang, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out["iou_con_loss"].cpu(), want_iou, rtol=1e-4, atol=1e-6)
    (out["lang_con_loss"] + out["iou_con_loss"]).backward()
    assert m.pc_proj.weight.grad is not None and m.text_proj.weight.grad is not None
    early = m(to_dev(grounding_inputs(21, epoch=3), cuda))       # constrast_module.py:54-56
    assert float(early["con_loss"]) == 0.0 and "lang_con_loss" not in early


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        C = importlib.import_module("3dvlp_b200.contrast")
        dev = torch.device("cuda", rank)
        torch.manual_seed(0)
        m = C.ContrastModule(CONFIG).to(dev)                        # same weights on every rank
        datas = [grounding_inputs(60 + r, B=2, epoch=60) for r in range(world)]
        out = m(to_dev(datas[rank], dev), gather=True)
        local = m(to_dev(datas[rank], dev), gather=False)
        # expected sentence term: this rank's sentences against the kept objects of every rank
        import torch.nn.functional as F
        BU = importlib.import_module("3dvlp_b200.box_util")
        LG = importlib.import_module("3dvlp_b200.loss_grounding")
        boxes, objs = [], []
        for d in datas:
            dd = to_dev(d, dev)
            boxes.append(F.normalize(m.pc_proj(dd["bbox_feature"]), dim=-1))
            objs.append(dd["objectness_scores"].max(2)[1].bool())
        all_box, all_obj = torch.cat(boxes).reshape(-1, 128), torch.cat(objs).reshape(-1)
        dd = to_dev(datas[rank], dev)
        gc, gs = LG.gt_boxes_from_labels(dd, CONFIG)
        ious, _ = BU.box3d_diou_pairs(dd["pred_center"], dd["pred_size"], gc, gs + 1e-2)
        obj = objs[rank]
        target = ((ious > 0.25) & obj[:, None]).float()
        B, K = obj.shape
        Lmax = gc.shape[1]
        text = F.normalize(m.text_proj(dd["lang_emb"].view(B, Lmax, -1)), dim=-1)
        want = 0.
        for b in range(B):
            for j in range(int(dd["lang_num"][b])):
                logp = torch.log_softmax(text[b, j] @ all_box[all_obj].t(), -1)
                tgt = torch.zeros(all_obj.numel(), device=dev)
                tgt[(rank * B + b) * K:(rank * B + b + 1) * K] = target[b, j]
                want = want + (-(logp * tgt[all_obj]).sum() / obj[b].sum() / 2)
        torch.testing.assert_close(out["lang_con_loss"], want / B, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(out["iou_con_loss"], local["iou_con_loss"])
        out["lang_con_loss"].backward()
        assert m.pc_proj.weight.grad is not None
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_contrast_all_gather_world2_nccl(cuda):
    """The north star's one collective over NCCL/NVLink: needs two GPUs (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
