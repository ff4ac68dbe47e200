from .bulldozer import ForestFireBulldozerEnv
from .advanced_bulldozer import AdvancedForestFireBulldozerEnv, BatchedAdvancedBulldozerEnv, MDP

__all__ = ["ForestFireBulldozerEnv", "AdvancedForestFireBulldozerEnv", "BatchedAdvancedBulldozerEnv", "MDP"]
