"""B200-native drop-in for the reference's ``models`` package: the same eight public names (``models/__init__.py:5-12``),
each backed by the sm_100a engine behind ``include/rfv.h`` instead of PyTorch layers."""
from . import base_flow as _base, rectified_flow as _rect, unet as _unet

UNet, count_parameters = _unet.UNet, _unet.count_parameters
BaseFlowModel, train_base_flow = _base.BaseFlowModel, _base.train_base_flow
RectifiedFlowModel = _rect.RectifiedFlowModel
generate_reflow_pairs, train_rectified_flow, iterative_reflow = (_rect.generate_reflow_pairs, _rect.train_rectified_flow,
                                                                 _rect.iterative_reflow)

__version__ = "0.1.0"
__all__ = ["UNet", "count_parameters", "BaseFlowModel", "train_base_flow", "RectifiedFlowModel", "generate_reflow_pairs",
           "train_rectified_flow", "iterative_reflow"]
