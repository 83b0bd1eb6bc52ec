"""``python -m tblup_b200.main <reference main.py arguments>``: run the UNMODIFIED reference main loop
(main.py:14-45 of ianwhale/tblup) with the B200 evaluators plugged in.  The reference checkout is found on
``sys.path`` or through ``TBLUP_REFERENCE=/path/to/tblup-checkout``."""
import os
import runpy
import sys


def main():
    ref = os.environ.get("TBLUP_REFERENCE")
    if ref and ref not in sys.path:
        sys.path.insert(0, ref)
    from .install import install
    try:
        import tblup
    except ImportError as e:
        raise SystemExit("tblup_b200.main: the reference package `tblup` is not importable; "
                         "set TBLUP_REFERENCE to its checkout") from e
    install(tblup)
    main_py = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(tblup.__file__))), "main.py")
    runpy.run_path(main_py, run_name="__main__")


if __name__ == "__main__":
    main()
