"""TEST INFRASTRUCTURE ONLY.  ctypes loaders for the oracle port (oracle/libfmrx_oracle.so) and, when it was built,
the unmodified reference behind its shim (oracle/_ref/libfmref.so) and the reference binary (oracle/_ref/fm_radio).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
from .port import Port, Chain, RdsDecoder, TAPS, load_port, build  # noqa: F401
from .ref import Ref, RefChain, load_ref, ref_available, ref_binary, run_ref_binary  # noqa: F401
