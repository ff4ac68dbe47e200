from .advanced_bulldozer import AdvancedForestFireBulldozerEnv, BatchedAdvancedBulldozerEnv, MDP

__all__ = ["AdvancedForestFireBulldozerEnv", "BatchedAdvancedBulldozerEnv", "MDP"]
