"""test_isaacgym_b200 -- B200-native (sm_100a) per-environment control laws behind the call
signatures of wp133716/test_isaacgym.

Modules mirror the reference files they replace:

=================================  ==========================================================
``controller6``                    ``common/controller6.py`` (cclvf2, euler2quaternion, CameraController)
``secondary_control_vecenv``       ``common/secondary_control_vecenv.py`` (SecondaryControl.servo_ext_pixel)
``servo_controller``               ``common/servo_controller.py`` / ``servo_controller_debug.py`` (scalar API)
``franka_cube_ik_osc``             ``examples/franka_cube_ik_osc.py`` control_ik / control_osc / orientation_error
``pd_control``                     joint PD torque law on dof_state -> dof_actuation_force (north-star entry)
``servo_step``                     fused ``test10_servo_vecenv.py:403-456`` step, in place on the root state
``sharding``                       env slices per GPU + statistics all-reduce
=================================  ==========================================================

Every law executes in ``libb200ctl.so`` (hand-written CUDA, ``csrc/``) through the C ABI in
``include/b200ctl.h``.  There is no CPU or eager fallback: the first call raises if the library is
missing or no CUDA device is visible.
"""
__version__ = "0.1.0"
