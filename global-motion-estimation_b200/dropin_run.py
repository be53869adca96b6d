"""Runs one of the reference's OWN scripts (results.py, bbme.py, "test scripts/motion_compensation.py") unchanged
against the B200 modules.

    python global-motion-estimation_b200/dropin_run.py /path/to/global_motion_estimation/results.py -v pan240.mp4 -f 3

The reference imports ``bbme``, ``motion`` and ``utils`` by bare name (results.py:1-2, bbme.py:9, motion.py:4-6), so
running its files directly would put the reference directory at sys.path[0] and shadow the drop-in modules.  This
launcher puts THIS directory first and executes the script with runpy under ``__main__``; the working directory is
left alone (results.py reads resources/videos/<name> and writes results/ relative to it, results.py:20-34).
"""
import os
import runpy
import sys


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:] = [here] + [p for p in sys.path if os.path.abspath(p or ".") not in (here, os.path.dirname(script))]
    for name in ("bbme", "motion", "utils"):
        sys.modules.pop(name, None)
    sys.argv = [script] + argv[2:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
