"""Runs bench.py under a watchdog: prints every thread's Python stack and exits if the run exceeds the limit (first argument,
seconds) -- a hung kernel then costs seconds of GPU time instead of the gpurun limit."""
import faulthandler
import runpy
import sys
limit = int(sys.argv[1])
faulthandler.dump_traceback_later(limit, exit=True)
sys.argv = ["bench.py"] + sys.argv[2:]
runpy.run_path("bench.py", run_name="__main__")
