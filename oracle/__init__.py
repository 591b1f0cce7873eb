"""CPU oracle for the hybrid-ODE hot path.  TEST INFRASTRUCTURE ONLY -- never imported by the product package.

* ``oracle.odeint``  restates torchdiffeq 0.2.2 (un-vendored third-party dependency of the reference).
* ``oracle.fields``  restates the reference's vector fields / decoder / masked SSE (``/root/reference/model.py``).
* ``oracle.refload`` imports the reference's own ``model.py`` unmodified (this container only) with
  ``oracle.odeint`` injected as ``torchdiffeq``; used to pin ``oracle.fields`` and to generate ``tests/golden``.

PARITY STATUS: the reference has no solver-level tests or golden vectors ("parity unpinned" by the reference);
the vector fields, decoder and loss ARE pinned bit-for-bit against the reference's own code (tests/golden).
"""
