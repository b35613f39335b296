"""CPU oracle for the per-environment control-law hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``test_isaacgym_b200/`` may import this
package: the product path is the sm_100a CUDA library behind ``include/b200ctl.h``
and fails loudly when that library is missing.  The only legitimate importers
are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, where the oracle is the checker or the timed CPU
baseline, never the thing shipped.

The reference (wp133716/test_isaacgym) is pure Python, so the oracle is a
numpy / scipy / torch-CPU restatement of its arithmetic, one module per kernel
family, each function citing the reference file:line it follows:

* ``oracle.servo``  -- family S: ``cclvf2`` -> ``world2pixel`` -> ``servo_ext_pixel``
  -> ``euler2quaternion`` -> root-state scatter (``common/controller6.py``,
  ``common/secondary_control_vecenv.py``, ``test10_servo_vecenv.py:403-456``).
* ``oracle.pd``     -- family P: joint PD torque law (fragments at
  ``examples/franka_cube_ik_osc.py:74-76``, ``examples/franka_osc.py:241``,
  ``examples/dof_controls.py:180-181``).
* ``oracle.franka`` -- family O: ``control_ik`` / ``control_osc`` /
  ``orientation_error`` (``examples/franka_cube_ik_osc.py:34-79``).

Pinning status (see DESIGN.md "Oracle"):

* S is pinned: the four known-answer vectors in the reference's ``__main__``
  blocks plus fixtures produced by importing the real reference modules in the
  build container (``tests/golden/gen_golden.py``, outputs committed under
  ``tests/golden/``).
* O is pinned to fixtures produced by AST-extracting the reference's own
  ``control_ik`` / ``control_osc`` / ``orientation_error`` and running them on
  seeded inputs (same script).  ``isaacgym.torch_utils`` (quat_mul,
  quat_conjugate) is an un-vendored, un-installable dependency: restated from
  the Hamilton product definition -- that part is "parity unpinned".
* P has NO reference function: "parity unpinned" for the law as a whole; the
  restatement is pinned only to the three reference fragments it reduces to.
"""
