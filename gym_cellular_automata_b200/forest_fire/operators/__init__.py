from .ca_alexandridis_cuda import PartiallyObservableForestFireCUDA
from .ca_windy_cuda import WindyForestFireCUDA
from .move_modify_cuda import ModifyCUDA, MoveCUDA, MoveModifyCUDA
from .repeat_ca_cuda import RepeatCACUDA

# drop-in names of the reference's JAX operators (forest_fire/operators/__init__.py:1-20)
PartiallyObservableForestFireJax = PartiallyObservableForestFireCUDA
MoveJax, ModifyJax, MoveModifyJax = MoveCUDA, ModifyCUDA, MoveModifyCUDA
RepeatCAJax = RepeatCACUDA
WindyForestFire = WindyForestFireCUDA  # v3 rule set (reference operators/ca_windy.py)

__all__ = ["WindyForestFireCUDA", "WindyForestFire", "PartiallyObservableForestFireCUDA", "MoveCUDA", "ModifyCUDA", "MoveModifyCUDA", "RepeatCACUDA",
           "PartiallyObservableForestFireJax", "MoveJax", "ModifyJax", "MoveModifyJax", "RepeatCAJax"]
