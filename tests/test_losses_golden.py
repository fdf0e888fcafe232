"""The loss rows of SURVEY.md 8(f) ranks 2-4 against tests/golden/reference_losses.npz -- outputs of the
REFERENCE's own loss_detection.py / loss_grounding.py / constrast_module.py (tests/golden/make_golden_losses.py).

CPU suite: (i) the oracle's loop restatements (oracle/boxes.py) are pinned to the reference; (ii) the host
logic of the product's loss modules runs on CPU with the two CUDA-only ops they call (nn_distance, the
pairwise box IoU) replaced by the oracle's restatements -- a test of the glue, values and gradients.
GPU suite: the product itself (CUDA kernels included) against the same fixture."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from common import det_tensor  # noqa: E402
from loss_cases import GROUNDING_CASES, MEAN_SIZES, clone_dict, detection_inputs, grounding_inputs, make_config  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_losses.npz"))


@pytest.fixture(scope="module")
def OB():
    from oracle import boxes
    return boxes


def _to(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def _np(v):
    return v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)


@pytest.fixture
def on_cpu(monkeypatch, pkg, oracle, OB):
    """Product loss modules with their two CUDA-only ops swapped for the oracle's restatements."""
    LD = importlib.import_module("3dvlp_b200.loss_detection")
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    C = importlib.import_module("3dvlp_b200.contrast")

    def pairs(pc, ps, gc, gs):
        B, K = pc.shape[:2]
        L = gc.shape[1]
        i, d = OB.box3d_diou_batch_tensor(pc[:, None].expand(B, L, K, 3).reshape(-1, 3), ps[:, None].expand(B, L, K, 3).reshape(-1, 3),
                                          gc[:, :, None].expand(B, L, K, 3).reshape(-1, 3), gs[:, :, None].expand(B, L, K, 3).reshape(-1, 3))
        return i.view(B, L, K), d.view(B, L, K)

    monkeypatch.setattr(LD, "nn_distance", oracle.nn_distance_torch)
    monkeypatch.setattr(LG, "box3d_diou_pairs", pairs)
    monkeypatch.setattr(LG, "box3d_diou_batch_tensor", OB.box3d_diou_batch_tensor)
    monkeypatch.setattr(C, "box3d_diou_pairs", pairs)
    return torch.device("cpu")


DET_GRADS = ("vote_xyz", "aggregated_vote_xyz", "objectness_scores", "heading_scores", "heading_residuals_normalized",
             "rois", "sem_cls_scores")


def _detection(dev, gold):
    LD = importlib.import_module("3dvlp_b200.loss_detection")
    cfg = make_config()
    d = _to(detection_inputs(51), dev)
    for k in DET_GRADS:
        d[k].requires_grad_(True)
    vote_loss = LD.compute_vote_loss(d)
    obj_loss, label, mask, assign = LD.compute_objectness_loss(d)
    assert assign.dtype == torch.int64 and label.dtype == torch.int64
    np.testing.assert_array_equal(_np(label), gold["det_objectness_label"])
    np.testing.assert_array_equal(_np(mask), gold["det_objectness_mask"])
    np.testing.assert_array_equal(_np(assign), gold["det_object_assignment"])
    assert 0 < int(label.sum()) < label.numel()
    d["objectness_label"], d["objectness_mask"], d["object_assignment"] = label, mask, assign
    rec = LD.recover_assigned_gt_bboxes(d, cfg, assign)
    names = ("gt_center", "heading_class_label", "heading_residual_label", "gt_heading", "gt_distance", "inside_label",
             "gt_centerness", "gt_bbox_size")
    for n, v in zip(names, rec):
        want = gold["det_rec_" + n]
        assert tuple(v.shape) == want.shape, n
        if v.dtype in (torch.int64, torch.int32):
            np.testing.assert_array_equal(_np(v), want, err_msg=n)
        else:
            np.testing.assert_allclose(_np(v), want, rtol=1e-5, atol=1e-6, err_msg=n)
    assert 0 < int(rec[5].sum()) < rec[5].numel()   # proposals inside and outside their assigned box
    hc, hr, dist, sem = LD.compute_box_and_sem_cls_loss(d, cfg)
    for n, v in (("vote_loss", vote_loss), ("objectness_loss", obj_loss), ("heading_cls_loss", hc), ("heading_reg_loss", hr),
                 ("size_distance_loss", dist), ("sem_cls_loss", sem)):
        np.testing.assert_allclose(_np(v), gold["det_" + n], rtol=1e-5, atol=1e-6, err_msg=n)
    for n in ("gt_assigned_center", "gt_assigned_heading_class", "gt_assigned_heading_residual", "gt_assigned_heading",
              "gt_assigned_distance", "gt_assigned_centerness"):
        np.testing.assert_allclose(_np(d[n]), gold["det_" + n], rtol=1e-5, atol=1e-6, err_msg=n)
    total = vote_loss + 0.5 * obj_loss + (0.1 * hc + hr + 0.1 * sem + 20 * dist)
    total.backward()
    for n in DET_GRADS:
        g = d[n].grad if d[n].grad is not None else torch.zeros_like(d[n])
        np.testing.assert_allclose(_np(g), gold["det_grad_" + n], rtol=1e-4, atol=1e-7, err_msg="grad " + n)
    # the one-call form used by a training step (lib/loss_helper/loss_joint.py:39-75)
    d2 = _to(detection_inputs(51), dev)
    out = LD.get_detection_loss(d2, cfg)
    want_box = 0.1 * gold["det_heading_cls_loss"] + gold["det_heading_reg_loss"] + 0.1 * gold["det_sem_cls_loss"] \
        + 20 * gold["det_size_distance_loss"]
    np.testing.assert_allclose(_np(out["box_loss"]), want_box, rtol=1e-5)
    assert 0.0 <= float(out["obj_acc"]) <= 1.0 and float(out["pos_ratio"]) > 0


def test_detection_losses_host_logic_vs_reference_golden(on_cpu, gold):
    _detection(on_cpu, gold)


@pytest.mark.gpu
def test_detection_losses_vs_reference_golden(pkg, cuda, gold):
    _detection(cuda, gold)


def _grounding(dev, gold, tag):
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    seed, kw_in, kw_fn = GROUNDING_CASES[tag]
    cfg = make_config()
    cfg.mean_size_arr_tensor = cfg.mean_size_arr_tensor.to(dev)
    grads = ["pred_center", "pred_size", "cluster_ref"] + (["pred_center_reg", "pred_size_reg"] if kw_fn.get("use_reg_head") else []) \
        + (["alpha"] if kw_fn.get("use_kl_loss") else [])
    d = _to(grounding_inputs(seed, **kw_in), dev)
    for k in grads:
        d[k].requires_grad_(True)
    dd, loss, preds, labels = LG.compute_diou_loss(d, cfg, **kw_fn)
    np.testing.assert_array_equal(_np(labels), gold[f"gr_{tag}_labels"])
    assert float(labels.sum()) >= 3
    np.testing.assert_allclose(_np(loss), gold[f"gr_{tag}_loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_np(preds), gold[f"gr_{tag}_preds"])
    scalars = ["diou_loss", "max_iou_rate_0.25", "max_iou_rate_0.5", "top_ind", "pred_iou_rate_0.25", "pred_iou_rate_0.5"] \
        + [f"class_iou_rate_{i}" for i in range(18)] + [f"class_size_{i}" for i in range(18)] \
        + [f"top_iou_rate_{i}" for i in range(1, 6)] + (["kl_loss"] if kw_fn.get("use_kl_loss") else [])
    for k in scalars:   # every key the reference solver reads (lib/joint/solver_3dvlp.py:651-661,687-697)
        assert k in dd, k
        np.testing.assert_allclose(_np(dd[k]), gold[f"gr_{tag}_{k}"], rtol=1e-5, atol=1e-6, err_msg=k)
    total = loss + dd["diou_loss"] + (dd["kl_loss"] if kw_fn.get("use_kl_loss") else 0.)
    total.backward()
    for n in grads:
        g = d[n].grad if d[n].grad is not None else torch.zeros_like(d[n])
        np.testing.assert_allclose(_np(g), gold[f"gr_{tag}_grad_{n}"], rtol=1e-4, atol=1e-7, err_msg="grad " + n)


@pytest.mark.parametrize("tag", sorted(GROUNDING_CASES))
def test_diou_loss_host_logic_vs_reference_golden(on_cpu, gold, tag):
    _grounding(on_cpu, gold, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(GROUNDING_CASES))
def test_diou_loss_vs_reference_golden(pkg, cuda, gold, tag):
    _grounding(cuda, gold, tag)


def test_diou_loss_istrain_may_be_a_device_tensor_or_a_host_value(on_cpu):
    LG = importlib.import_module("3dvlp_b200.loss_grounding")
    outs = []
    for istrain in (torch.tensor([1]), [1], 1):
        d = grounding_inputs(31, epoch=10, istrain=1, rnd=0.3)
        d["istrain"] = istrain
        outs.append(LG.compute_diou_loss(d, make_config())[1])
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("tag", ["early_masked", "early_plain", "late_eval", "early_reg", "late_reg", "no_reference"])
def test_oracle_diou_loss_loop_vs_reference_golden(OB, gold, tag):
    """oracle/boxes.py compute_diou_loss_loop (the loop restatement the round-1 tests compared with) is
    itself pinned to the reference function."""
    seed, kw_in, kw_fn = GROUNDING_CASES[tag]
    d = grounding_inputs(seed, **kw_in)
    loss, diou, labels, r25, r5 = OB.compute_diou_loss_loop(d, MEAN_SIZES, **kw_fn)
    np.testing.assert_array_equal(labels.numpy(), gold[f"gr_{tag}_labels"])
    np.testing.assert_allclose(_np(loss), gold[f"gr_{tag}_loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_np(diou), gold[f"gr_{tag}_diou_loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r25, gold[f"gr_{tag}_max_iou_rate_0.25"], rtol=1e-6)
    np.testing.assert_allclose(r5, gold[f"gr_{tag}_max_iou_rate_0.5"], rtol=1e-6)


def _contrast_weights():
    return {"pc_proj.weight": det_tensor("pc_proj", (128, 128), 0.15), "text_proj.weight": det_tensor("text_proj", (128, 128), 0.15),
            "pc_proj_iou.0.weight": det_tensor("pc_proj_iou", (128, 128), 0.15)}


def _contrast(dev, gold):
    C = importlib.import_module("3dvlp_b200.contrast")
    cfg = make_config()
    cfg.mean_size_arr_tensor = cfg.mean_size_arr_tensor.to(dev)
    m = C.ContrastModule(cfg)
    assert sorted(m.state_dict().keys()) == list(gold["con_state_keys"])
    m.load_state_dict(dict(_contrast_weights(), **{"nce_loss.tau": m.nce_loss.tau.detach().clone()}))
    m = m.to(dev)
    for tag, seed in (("a", 61), ("b", 62)):
        d = _to(grounding_inputs(seed, B=4 if tag == "b" else 3, epoch=60), dev)
        d["bbox_feature"].requires_grad_(True); d["lang_emb"].requires_grad_(True)
        m.zero_grad()
        r = m(d)
        np.testing.assert_allclose(_np(r["lang_con_loss"]), gold[f"con_{tag}_lang_con_loss"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(_np(r["iou_con_loss"]), gold[f"con_{tag}_iou_con_loss"], rtol=1e-4, atol=1e-6)
        (r["lang_con_loss"] + 2.0 * r["iou_con_loss"]).backward()
        np.testing.assert_allclose(_np(d["bbox_feature"].grad), gold[f"con_{tag}_grad_bbox_feature"], rtol=1e-3, atol=1e-7)
        np.testing.assert_allclose(_np(d["lang_emb"].grad), gold[f"con_{tag}_grad_lang_emb"], rtol=1e-3, atol=1e-7)
        for n, p in m.named_parameters():
            if f"con_{tag}_grad_{n}" in gold.files:
                np.testing.assert_allclose(_np(p.grad), gold[f"con_{tag}_grad_{n}"], rtol=1e-3, atol=1e-6, err_msg=n)
    early = m(_to(grounding_inputs(61, epoch=3), dev))       # constrast_module.py:54-56
    assert float(early["con_loss"]) == float(gold["con_early_con_loss"].reshape(-1)[0]) == 0.0 and "lang_con_loss" not in early


def test_contrast_module_host_logic_vs_reference_golden(on_cpu, gold):
    _contrast(on_cpu, gold)


@pytest.mark.gpu
def test_contrast_module_vs_reference_golden(pkg, cuda, gold):
    _contrast(cuda, gold)


def test_oracle_contrast_loop_vs_reference_golden(OB, gold):
    w = _contrast_weights()
    for tag, seed in (("a", 61), ("b", 62)):
        d = grounding_inputs(seed, B=4 if tag == "b" else 3, epoch=60)
        lang, iou = OB.contrast_loop(d, MEAN_SIZES, w["pc_proj.weight"], w["text_proj.weight"], w["pc_proj_iou.0.weight"])
        np.testing.assert_allclose(_np(lang), gold[f"con_{tag}_lang_con_loss"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(_np(iou), gold[f"con_{tag}_iou_con_loss"], rtol=1e-4, atol=1e-6)
