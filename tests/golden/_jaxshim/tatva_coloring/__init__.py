"""Stand-in for the PyPI package `tatva-coloring` (not in the image, no source available).

Delegates to the reference's own in-tree predecessor, tatva/sparse/_coloring.py, loaded
by file path so that `tatva.sparse` (which imports *this* package) is not re-entered.
"""
import importlib.util as _ilu
import os as _os

_p = _os.path.join(_os.environ.get("TATVA_REFERENCE", "/root/reference"), "tatva", "sparse", "_coloring.py")
_spec = _ilu.spec_from_file_location("_tatva_ref_coloring", _p)
_m = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_m)

distance2_colors = _m.distance2_colors
distance2_color_and_seeds = _m.distance2_color_and_seeds
