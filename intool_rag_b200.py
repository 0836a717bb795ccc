"""Import shim: the product package lives in the directory ``intool-rag_b200/`` (the name the
task fixes), which is not a valid Python identifier.  ``import intool_rag_b200`` loads that
directory as a regular package, so ``intool_rag_b200.faiss`` etc. resolve to files inside it."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "intool-rag_b200")
_spec = _u.spec_from_file_location("intool_rag_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["intool_rag_b200"] = _mod
_spec.loader.exec_module(_mod)
