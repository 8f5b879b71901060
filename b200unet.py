"""Import shim: ``import b200unet`` -> the package directory
``adaptive-depth-u-net-for-image-super-resolution-segmentation_b200/`` (whose
mandated name contains hyphens and so cannot be written in an import statement)."""
import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().parent / "adaptive-depth-u-net-for-image-super-resolution-segmentation_b200"
_spec = importlib.util.spec_from_file_location("b200unet", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200unet"] = _mod
_spec.loader.exec_module(_mod)
