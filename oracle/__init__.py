"""CPU oracle for the sampling hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package under
``image-inpainting-..._b200/`` never does.  See ``oracle/unet.py`` for the
reference file:line map and the pinning status of each piece.
"""
