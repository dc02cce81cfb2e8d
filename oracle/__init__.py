"""oracle/ -- TEST INFRASTRUCTURE (CPU checker), never imported by the product package.

``oracle.port``  : ctypes view of oracle/liboracle.so (plain-C restatement, goetia_oracle.c)
``oracle.ref``   : ctypes view of oracle/_ref/libgoetia_ref.so (the unmodified reference,
                   compiled from /root/reference by oracle/Makefile), or None when absent.
"""
from .binding import Port, Ref, build_oracle, build_ref, have_ref, synth_reads  # noqa: F401
