"""dl_biomass_b200 -- B200-native (sm_100a) PointNet++ set-abstraction path of cczls1991/DL_Biomass."""
from .data import Batch, Data, synthetic_cloud, synthetic_clouds  # noqa: F401

__all__ = ["Batch", "Data", "synthetic_cloud", "synthetic_clouds", "Net", "SAModule", "GlobalSAModule", "MLP"]


def __getattr__(name):  # the model classes pull in torch.library registration; import them lazily
    if name in ("Net", "SAModule", "GlobalSAModule", "MLP", "PointConv"):
        from . import pointnet2_regressor as m
        return getattr(m, name)
    raise AttributeError(name)
