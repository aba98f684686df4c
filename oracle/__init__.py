"""CPU oracle for the env-step hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Float64 NumPy restatement of the reference's algorithms (each function cites the reference
file:line it follows, or the pinned third-party module whose published algorithm it restates).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import anything from this package.  The product package
``olympics_mujoco_b200`` never imports it and fails loudly when its CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):
* kinematics / COM / COM-velocity (MuJoCo 2.3.6 C engine, not in /root/reference, not installable
  here): **parity unpinned by the reference** -- pinned only by analytic known-answer checks and a
  second, independently written C restatement (``oracle/c``).
* trajectory handling, phase clocks, PPO returns, discriminator networks: pinned against the
  reference's own Python modules imported in the authoring container; fixtures under
  ``tests/golden`` (generator: ``tools/gen_golden.py``).
"""
