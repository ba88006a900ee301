"""Drop-in module name of the reference (/root/reference/pointnet2_regressor.py): ``main.py``,
``hyperparameter_tuning.py`` and ``point_density_effect.py`` do ``from pointnet2_regressor import Net``."""
from dl_biomass_b200.pointnet2_regressor import MLP, GlobalSAModule, Net, PointConv, SAModule  # noqa: F401
