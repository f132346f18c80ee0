"""CPU oracle for the Dr.Jit-Core primitive path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this package. The product
(``drjit_b200``) never does.

``oracle.capi``  : numpy front-end of ``oracle.c`` (the restatement).
``oracle.ref``   : ctypes front-end of the UNMODIFIED reference library built by
                   ``oracle/ref_build/Makefile`` into ``oracle/_ref/`` (may be absent).
"""
