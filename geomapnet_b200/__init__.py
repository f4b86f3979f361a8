"""geomapnet_b200 -- Blackwell-native MapNet/PoseNet training hot path.

Package layout mirrors the reference's import paths for THIS path only
(``models.posenet``, ``common.criterion``, ``common.optimizer``) so that
``scripts/train.py`` resolves them when this directory is first on sys.path
(see INTEGRATION.md); everything heavy lives in csrc/ behind the C ABI declared
in include/mapnet_b200.h.
"""
__version__ = "0.1.0"
