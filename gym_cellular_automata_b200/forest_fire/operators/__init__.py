from .ca_alexandridis_cuda import PartiallyObservableForestFireCUDA
from .move_modify_cuda import ModifyCUDA, MoveCUDA, MoveModifyCUDA
from .repeat_ca_cuda import RepeatCACUDA

# drop-in names of the reference's JAX operators (forest_fire/operators/__init__.py:1-20)
PartiallyObservableForestFireJax = PartiallyObservableForestFireCUDA
MoveJax, ModifyJax, MoveModifyJax = MoveCUDA, ModifyCUDA, MoveModifyCUDA
RepeatCAJax = RepeatCACUDA

__all__ = ["PartiallyObservableForestFireCUDA", "MoveCUDA", "ModifyCUDA", "MoveModifyCUDA", "RepeatCACUDA",
           "PartiallyObservableForestFireJax", "MoveJax", "ModifyJax", "MoveModifyJax", "RepeatCAJax"]
