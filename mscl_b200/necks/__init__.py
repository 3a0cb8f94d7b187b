from .moco_necks import BaseMoCo, TPNMoCo

__all__ = ["BaseMoCo", "TPNMoCo"]
