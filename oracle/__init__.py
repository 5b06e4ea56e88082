"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Garment4D hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
``garment4d_b200`` never imports it.

* :mod:`oracle.pointnet2` -- numpy front-end over ``liboracle.so``, the plain-C
  restatement of the reference's pointnet2 CUDA kernels (``pointnet2_oracle.c``).
* :mod:`oracle.lbs` -- numpy fp32 restatement of ``smplx/smplx/lbs.py``.
* :mod:`oracle.refgpu` -- ctypes front-end over ``oracle/_ref/libpointnet2_ref.so``:
  the reference's own ``.cu`` files compiled unmodified (``build_ref.sh``); needs a GPU.
"""
