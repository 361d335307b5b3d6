"""Import alias: the package directory is named `nemotron-speech.cpp_b200` (not a Python identifier),
so `import nsb200` loads it under the module name `nemotron_speech_cpp_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nemotron-speech.cpp_b200")
_NAME = "nemotron_speech_cpp_b200"
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: getattr(_mod, k) for k in dir(_mod) if not k.startswith("__")})
