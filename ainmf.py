"""Import alias: the package directory is `audio-inpainting_b200/` (not a valid module name); `import ainmf` loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "audio-inpainting_b200")
_spec = importlib.util.spec_from_file_location("ainmf", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ainmf"] = _mod
_spec.loader.exec_module(_mod)
