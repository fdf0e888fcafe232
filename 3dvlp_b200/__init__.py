"""3dvlp_b200 -- B200-native (sm_100a) PointNet++ op suite + nn_distance for 3DVLP's detector backbone.

The package name starts with a digit, so import it with
``importlib.import_module("3dvlp_b200")`` (see INTEGRATION.md for the one-line drop-in).

Public surface = the reference's (file:line in each module docstring):
    _ext                 the nine functions of the pybind module ``pointnet2._ext``
    pointnet2_utils      furthest_point_sample, gather_operation, three_nn, three_interpolate,
                         grouping_operation, ball_query, QueryAndGroup, GroupAll
    pointnet2_modules    PointnetSAModuleVotes, PointnetFPModule (+ MSG / LFP variants)
    pytorch_utils        SharedMLP, BNMomentumScheduler, ...
    nn_distance          nn_distance, huber_loss, knn_distance
    backbone             Pointnet2Backbone
"""
import sys as _sys

from . import _lib  # noqa: F401
from ._lib import get_option, get_stat, set_option, set_option_thread, thread_options  # noqa: F401
from . import _ext, backbone, mlp_train, nn_distance as _nn_distance_mod, pointnet2_modules, pointnet2_utils, pytorch_utils, voting, loss_detection, box_util, loss_grounding, contrast, proposal  # noqa: F401
from .backbone import Pointnet2Backbone  # noqa: F401
from .mlp_train import mlp_backend  # noqa: F401
from .graph import GraphedBackbone, GraphedTrainStep  # noqa: F401
from .nn_distance import huber_loss, knn_distance, nn_distance  # noqa: F401
from .pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes  # noqa: F401
from .voting import VoteNetFrontEnd, VotingModule  # noqa: F401
from .contrast import ContrastModule  # noqa: F401
from .proposal import ProposalModule, StandardROIHeads, VoteNetDetector  # noqa: F401

__all__ = ["Pointnet2Backbone", "GraphedBackbone", "GraphedTrainStep", "VotingModule", "VoteNetFrontEnd", "ProposalModule", "StandardROIHeads", "VoteNetDetector", "PointnetSAModuleVotes", "PointnetFPModule", "nn_distance", "huber_loss",
           "knn_distance", "pointnet2_utils", "pointnet2_modules", "pytorch_utils", "install_dropin"]


def install_dropin():
    """Register this package under every import path the reference's callers use:
    ``pointnet2._ext``, ``pointnet2_utils``, ``pytorch_utils``, ``pointnet2_modules``,
    ``lib.pointnet2.*`` and ``utils.nn_distance``.  After this call the reference's model code
    (e.g. models/base_module/backbone_module.py:8) imports the B200 implementation unchanged."""
    import types
    pkg = _sys.modules[__name__]
    mods = {
        "pointnet2._ext": _ext,
        "pointnet2_utils": pointnet2_utils,
        "pytorch_utils": pytorch_utils,
        "pointnet2_modules": pointnet2_modules,
        "lib.pointnet2.pointnet2_utils": pointnet2_utils,
        "lib.pointnet2.pytorch_utils": pytorch_utils,
        "lib.pointnet2.pointnet2_modules": pointnet2_modules,
        "utils.nn_distance": _nn_distance_mod,
        "models.base_module.voting_module": voting,
        "models.base_module.backbone_module": backbone,
        "models.proposal_module.proposal_module_fcos": proposal,
        "models.proposal_module.ROI_heads.roi_heads": proposal,
    }
    import importlib
    for parent in ("pointnet2", "lib", "lib.pointnet2", "utils", "models", "models.base_module",
                   "models.proposal_module", "models.proposal_module.ROI_heads"):
        if parent in _sys.modules:
            continue
        try:  # keep the caller's real packages (the reference's lib/ and utils/ are namespace packages)
            importlib.import_module(parent)
        except Exception:
            m = types.ModuleType(parent)
            m.__path__ = []
            _sys.modules[parent] = m
    for name, mod in mods.items():
        _sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_sys.modules[parent], leaf, mod)
    if not hasattr(_sys.modules["lib"], "pointnet2"):
        setattr(_sys.modules["lib"], "pointnet2", _sys.modules["lib.pointnet2"])
    return pkg
