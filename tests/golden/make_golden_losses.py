"""Golden vectors for the loss rows of SURVEY.md 8(f) ranks 2-4, produced by the REFERENCE's own
functions imported unmodified from /root/reference (build container only):

    python tests/golden/make_golden_losses.py

* lib/loss_helper/loss_detection.py: compute_vote_loss, compute_objectness_loss,
  recover_assigned_gt_bboxes, compute_box_loss, compute_box_and_sem_cls_loss  (values + gradients)
* lib/loss_helper/loss_grounding.py: compute_diou_loss in eight configurations (label smoothing on/off,
  objectness masking, regression head, KL head, no_reference, debug statistics)  (values + gradients)
* models/constrast_module/constrast_module.py: ContrastModule.forward  (values + gradients)

What the import needs that this image lacks is registered as a stand-in BEFORE the import; none of it
computes a fixture value except the last one, which is stated where it is used:
  easydict          attribute dict (pulled in by lib/configs/config.py)
  torch_scatter     imported at the top of loss_grounding.py; only compute_attr_loss calls it (not here)
  plyfile, trimesh, matplotlib   file-IO / plotting libraries imported at the top of utils/pc_utils.py;
                    never called here
  Tensor.cuda() / torch.cuda.FloatTensor   the loss code hard-codes .cuda();
                    mapped to the CPU tensor itself so the reference arithmetic runs here
  pytorch3d.ops.iou_box3d.box3d_overlap    ContrastModule calls it on corner boxes it has just built
                    axis-aligned from centre + size (constrast_module.py:9-16,99-105); the stand-in
                    returns the closed-form volume / IoU of those axis-aligned boxes.  pytorch3d is not
                    vendored or pinned by the reference, so this one call stays "parity unpinned"; the
                    loops, masks, projections and NCE losses around it are the reference's own code.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)

from common import det_tensor  # noqa: E402
from loss_cases import GROUNDING_CASES, clone_dict, detection_inputs, grounding_inputs, make_config  # noqa: E402


def install_stand_ins():
    class EasyDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    m = types.ModuleType("easydict"); m.EasyDict = EasyDict
    sys.modules["easydict"] = m

    ts = types.ModuleType("torch_scatter")

    def scatter(*a, **k):  # never reached by the functions the fixture covers
        raise RuntimeError("torch_scatter stand-in: not available in this container")
    ts.scatter = scatter
    sys.modules["torch_scatter"] = ts

    def box3d_overlap(boxes1, boxes2, eps=1e-4):
        """(N,8,3),(M,8,3) corner boxes that are axis-aligned by construction -> (vol (N,M), iou (N,M))."""
        lo1, hi1 = boxes1.min(1)[0][:, None], boxes1.max(1)[0][:, None]
        lo2, hi2 = boxes2.min(1)[0][None], boxes2.max(1)[0][None]
        inter = torch.clamp(torch.min(hi1, hi2) - torch.max(lo1, lo2), min=0).prod(-1)
        v1, v2 = (hi1 - lo1).prod(-1), (hi2 - lo2).prod(-1)
        return inter, inter / (v1 + v2 - inter)
    for name in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.iou_box3d"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d.ops.iou_box3d"].box3d_overlap = box3d_overlap
    sys.modules["pytorch3d"].ops = sys.modules["pytorch3d.ops"]
    sys.modules["pytorch3d.ops"].iou_box3d = sys.modules["pytorch3d.ops.iou_box3d"]

    # mesh / point-cloud file IO pulled in by utils/pc_utils.py (through utils/utils_fn.py); never called here
    ply = types.ModuleType("plyfile"); ply.PlyData = ply.PlyElement = None
    sys.modules.setdefault("plyfile", ply)
    for name in ("trimesh", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = types.ModuleType(name)
            if "." in name:
                setattr(sys.modules[name.rpartition(".")[0]], name.rpartition(".")[2], sys.modules[name])
    if not hasattr(sys.modules["matplotlib.pyplot"], "cm"):  # a default argument of a plotting helper reads pyplot.cm.jet
        sys.modules["matplotlib.pyplot"].cm = types.SimpleNamespace(jet=None)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor


def put(out, key, v):
    if torch.is_tensor(v):
        v = v.detach().cpu().numpy()
    out[key] = np.asarray(v)


def detection(out):
    import lib.loss_helper.loss_detection as LD
    from data.scannet.model_util_scannet import ScannetDatasetConfig
    cfg = make_config(ScannetDatasetConfig)
    grads = ("vote_xyz", "aggregated_vote_xyz", "objectness_scores", "heading_scores", "heading_residuals_normalized",
             "rois", "sem_cls_scores")
    d = clone_dict(detection_inputs(51), grad=grads)
    vote_loss = LD.compute_vote_loss(d)
    obj_loss, obj_label, obj_mask, assignment = LD.compute_objectness_loss(d)
    d["objectness_label"], d["objectness_mask"], d["object_assignment"] = obj_label, obj_mask, assignment
    rec = LD.recover_assigned_gt_bboxes(d, cfg, assignment)
    names = ("gt_center", "heading_class_label", "heading_residual_label", "gt_heading", "gt_distance", "inside_label",
             "gt_centerness", "gt_bbox_size")
    for n, v in zip(names, rec):
        put(out, "det_rec_" + n, v)
    hc, hr, dist, sem = LD.compute_box_and_sem_cls_loss(d, cfg)
    for n, v in (("vote_loss", vote_loss), ("objectness_loss", obj_loss), ("objectness_label", obj_label),
                 ("objectness_mask", obj_mask), ("object_assignment", assignment), ("heading_cls_loss", hc),
                 ("heading_reg_loss", hr), ("size_distance_loss", dist), ("sem_cls_loss", sem)):
        put(out, "det_" + n, v)
    for n in ("gt_assigned_center", "gt_assigned_heading_class", "gt_assigned_heading_residual", "gt_assigned_heading",
              "gt_assigned_distance", "gt_assigned_centerness"):
        put(out, "det_" + n, d[n])
    # the combination of lib/loss_helper/loss_joint.py:52-55 plus vote and objectness terms
    total = vote_loss + 0.5 * obj_loss + (0.1 * hc + hr + 0.1 * sem + 20 * dist)
    total.backward()
    for n in grads:
        put(out, "det_grad_" + n, d[n].grad if d[n].grad is not None else torch.zeros_like(d[n]))


def grounding(out):
    import lib.loss_helper.loss_grounding as LG
    from data.scannet.model_util_scannet import ScannetDatasetConfig
    cfg = make_config(ScannetDatasetConfig)
    for tag, (seed, kw_in, kw_fn) in GROUNDING_CASES.items():
        grads = ["pred_center", "pred_size", "cluster_ref"] + (["pred_center_reg", "pred_size_reg"] if kw_fn.get("use_reg_head") else []) \
            + (["alpha"] if kw_fn.get("use_kl_loss") else [])
        d = clone_dict(grounding_inputs(seed, **kw_in), grad=grads)
        dd, loss, preds, labels = LG.compute_diou_loss(d, cfg, **kw_fn)
        put(out, f"gr_{tag}_loss", loss); put(out, f"gr_{tag}_labels", labels); put(out, f"gr_{tag}_preds", preds)
        scalars = ["diou_loss", "max_iou_rate_0.25", "max_iou_rate_0.5", "top_ind", "pred_iou_rate_0.25", "pred_iou_rate_0.5"] \
            + [f"class_iou_rate_{i}" for i in range(18)] + [f"class_size_{i}" for i in range(18)] \
            + [f"top_iou_rate_{i}" for i in range(1, 6)] + (["kl_loss"] if kw_fn.get("use_kl_loss") else [])
        for k in scalars:
            put(out, f"gr_{tag}_{k}", dd[k])
        total = loss + dd["diou_loss"] + (dd["kl_loss"] if kw_fn.get("use_kl_loss") else 0.)
        total.backward()
        for n in grads:
            put(out, f"gr_{tag}_grad_{n}", d[n].grad if d[n].grad is not None else torch.zeros_like(d[n]))


def contrast(out):
    from models.constrast_module.constrast_module import ContrastModule
    from data.scannet.model_util_scannet import ScannetDatasetConfig
    cfg = make_config(ScannetDatasetConfig)
    m = ContrastModule(cfg)
    sd = {"pc_proj.weight": det_tensor("pc_proj", (128, 128), 0.15), "text_proj.weight": det_tensor("text_proj", (128, 128), 0.15),
          "pc_proj_iou.0.weight": det_tensor("pc_proj_iou", (128, 128), 0.15), "nce_loss.tau": m.nce_loss.tau.detach().clone()}
    m.load_state_dict(sd)
    put(out, "con_state_keys", np.array(sorted(m.state_dict().keys())))
    for tag, seed in (("a", 61), ("b", 62)):
        d = clone_dict(grounding_inputs(seed, B=4 if tag == "b" else 3, epoch=60), grad=("bbox_feature", "lang_emb"))
        m.zero_grad()
        r = m(d)
        put(out, f"con_{tag}_lang_con_loss", r["lang_con_loss"]); put(out, f"con_{tag}_iou_con_loss", r["iou_con_loss"])
        (r["lang_con_loss"] + 2.0 * r["iou_con_loss"]).backward()
        put(out, f"con_{tag}_grad_bbox_feature", d["bbox_feature"].grad); put(out, f"con_{tag}_grad_lang_emb", d["lang_emb"].grad)
        for n, p in m.named_parameters():
            if p.grad is not None:
                put(out, f"con_{tag}_grad_{n}", p.grad)
    early = m(clone_dict(grounding_inputs(61, epoch=3)))
    put(out, "con_early_con_loss", early["con_loss"])


def main():
    torch.set_num_threads(1)
    install_stand_ins()
    sys.path.insert(0, REF)
    out = {}
    detection(out)
    grounding(out)
    contrast(out)
    np.savez_compressed(os.path.join(HERE, "reference_losses.npz"), **out)
    print("wrote", len(out), "arrays")
    for k in ("det_vote_loss", "det_objectness_loss", "det_size_distance_loss", "gr_early_masked_loss", "gr_debug_top_ind",
              "con_a_lang_con_loss", "con_a_iou_con_loss"):
        print(" ", k, out[k])


if __name__ == "__main__":
    main()
