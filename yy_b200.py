"""Import shim: the product package lives in ``yinyang-game-alphazero_b200/`` (hyphenated, as the project
layout prescribes), which Python cannot import by name.  Importing this module registers it as
``yinyang_game_alphazero_b200``."""
import importlib.util
import os
import sys

_NAME = "yinyang_game_alphazero_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "yinyang-game-alphazero_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

package = sys.modules[_NAME]
