"""Run-time override of the library's diagnostic switches for the probe / A-B tools.

The library reads FLOWCORR_* once per process; the tools flip switches between measurements through
``fc_tunable_set``.  FLOWCORR_PROBE (stage probes) exists only in a library built with
``make -C flow_supervisor_b200/csrc EXTRA=-DFC_PROBES OUT=../libflowcorr_probes.so`` -- point FLOWCORR_LIB at it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flow_supervisor_b200 import _lib  # noqa: E402

NAMES = {"FLOWCORR_PROBE": "probe", "FLOWCORR_BUILD_SCHED": "build_sched", "FLOWCORR_BUILD_STAGES": "build_stages",
         "FLOWCORR_BUILD_EPI_WARPS": "build_epi_warps", "FLOWCORR_NO_FUSE": "no_fuse", "FLOWCORR_L2_FETCH": "l2_fetch"}


def set_switch(env_name: str, value) -> None:
    _lib.check(_lib.load().fc_tunable_set(NAMES[env_name].encode(), int(value)), f"fc_tunable_set({env_name})")
