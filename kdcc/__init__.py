"""Import alias: `import kdcc` loads the package that lives in the (non-identifier) directory
`knowledge-distillation-by-replacing-cheap-conv_b200/` and registers it as `kdcc`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "knowledge-distillation-by-replacing-cheap-conv_b200")


def _load():
    spec = importlib.util.spec_from_file_location("kdcc", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["kdcc"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
