"""Scenario plugin loader (reference: multiagent/scenarios/__init__.py:5-7).

`load("simple_spread.py")` returns the scenario *module* (the caller does `.Scenario()`,
make_env.py:36).  The reference uses `imp.load_source`, removed in Python 3.12; this loader
accepts the same argument (file name with or without `.py`, or a path to a scenario file)."""
import importlib
import importlib.util
import os.path as osp


def load(name):
    base = osp.basename(name)
    stem = base[:-3] if base.endswith(".py") else base
    pathname = osp.join(osp.dirname(__file__), stem + ".py")
    if osp.dirname(name) == "" and osp.exists(pathname):
        return importlib.import_module(__name__ + "." + stem)
    if not osp.exists(name):
        raise FileNotFoundError("no scenario file %r (built-ins live in %s)" % (name, osp.dirname(__file__)))
    spec = importlib.util.spec_from_file_location("mpe_b200_user_scenario_" + stem, name)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
