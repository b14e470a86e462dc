"""Runs one of the reference's own scripts (e.g. code/Raindrop.py) UNMODIFIED against this implementation:

    python -m raindrop_b200.launch /path/to/Raindrop/code/Raindrop.py --dataset P19 ...

`code/Raindrop.py:19` does `from models_rd import *`.  A plain `PYTHONPATH=<this repo> python Raindrop.py`
does NOT select this implementation: Python puts the script's own directory (`code/`, which holds the
reference's `models_rd.py`) at `sys.path[0]`, ahead of PYTHONPATH.  This launcher therefore
  1. pre-registers `raindrop_b200.models_rd` as `sys.modules["models_rd"]` (an import never consults the path
     for a module that is already loaded) and puts this repository's root first on `sys.path`,
  2. appends the script's directory so its sibling modules (`utils_rd`, ...) still import,
  3. changes into the script's directory (the reference uses relative data paths, code/Raindrop.py:74,163-170),
  4. executes the script as `__main__` with the remaining command-line arguments.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m raindrop_b200.launch <script.py> [script args...]")
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit("raindrop_b200.launch: no such script: %s" % script)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script_dir = os.path.dirname(script)
    sys.path[:] = [root] + [p for p in sys.path if os.path.abspath(p or ".") not in (root, script_dir)] + [script_dir]
    import raindrop_b200.models_rd as ours
    sys.modules["models_rd"] = ours
    os.chdir(script_dir)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
