"""Import alias: the package directory is named after the upstream project
(`predictive-multi-agent-framework_b200/`), which is not a valid Python identifier;
`import pmaf_b200` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "predictive-multi-agent-framework_b200")
_spec = importlib.util.spec_from_file_location(
    "pmaf_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pmaf_b200"] = _mod
_spec.loader.exec_module(_mod)
