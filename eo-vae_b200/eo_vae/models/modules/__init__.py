from .consistency_loss import EOConsistencyLoss
from .distributions import DiagonalGaussianDistribution
from .dynamic_conv import DynamicConv, DynamicConv_decoder

__all__ = ['EOConsistencyLoss', 'DiagonalGaussianDistribution', 'DynamicConv', 'DynamicConv_decoder']
