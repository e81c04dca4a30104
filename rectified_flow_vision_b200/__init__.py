"""B200-native drop-in for the reference's ``models`` package (``models/__init__.py:5-12``): same eight names."""
from .unet import UNet, count_parameters
from .base_flow import BaseFlowModel, train_base_flow
from .rectified_flow import (RectifiedFlowModel, generate_reflow_pairs, train_rectified_flow, iterative_reflow)

__all__ = ['UNet', 'count_parameters', 'BaseFlowModel', 'train_base_flow', 'RectifiedFlowModel',
           'generate_reflow_pairs', 'train_rectified_flow', 'iterative_reflow']
