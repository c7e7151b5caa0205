"""Debug helper: run smoke() with a watchdog that dumps the Python stack if a kernel hangs."""
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG_S", 50)), exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

g.smoke()
print("smoke finished")
