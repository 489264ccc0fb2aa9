"""svnet_b200 -- B200-native (sm_100a) implementation of SVNet's inference hot path behind the
reference's own nn.Module API.  ``import svnet_b200 as models`` mirrors the reference's
``import models`` for the SV classes (models/__init__.py:7-10)."""
from . import evalutil, sv_layers, sv_util
from .sv_layers import Conv1d, Linear, SV_STNkd, SVBlock, SVFuse, Vector2Scalar, VectorBN
from .sv_util import get_graph_feature, get_graph_feature_cross, get_graph_feature_sv, knn, svcat, svpool
from .sv_dgcnn_cls import SV_DGCNN_CLS
from .sv_dgcnn_partseg import SV_DGCNN_PSEG
from .sv_pointnet_cls import SV_PointNet_CLS
from .sv_pointnet_partseg import SV_PointNet_PSEG

__all__ = ["SV_DGCNN_CLS", "SV_DGCNN_PSEG", "SV_PointNet_CLS", "SV_PointNet_PSEG", "SVBlock", "SVFuse", "SV_STNkd", "Vector2Scalar", "VectorBN", "Linear", "Conv1d",
           "knn", "get_graph_feature", "get_graph_feature_cross", "get_graph_feature_sv", "svpool", "svcat",
           "sv_layers", "sv_util"]
from .graph import GraphedForward  # noqa: E402,F401
from .native_model import NativeModel  # noqa: E402,F401
