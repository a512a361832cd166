"""Loader shim: the package directory is named after the upstream repo (`nejm-brain-to-text_b200`),
which is not a valid Python identifier; this registers it as `nejm_brain_to_text_b200`."""
import importlib.util
import os
import sys

NAME = "nejm_brain_to_text_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "nejm-brain-to-text_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(NAME, None)
        raise
    return mod


def submodule(name):
    load()
    return importlib.import_module(f"{NAME}.{name}")
