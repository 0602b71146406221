#!/bin/bash
# Round 2, last GPU calls.  `density`: b2Fixture::SetDensity on the device (new parity test, the C++ mirror's program), then the
# edit regression tests that share k_body_commands.  `joints`: gear and mouse joints created on running worlds (host-side
# constructors over read-back state), the C++ program with its late mouse joint, then the edit test that creates plain joints.
mkdir -p gpurun_out
if [ "$1" = "mirror" ]; then  # the C++ mirror's three GPU programs (new getters: GetMass, flags, GetWorldPoint, fixture getters)
  timeout 40 python -m pytest tests/test_cpp_api.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2N_mirror.log
  exit 0
fi
if [ "$1" = "joints" ]; then
  timeout 45 python -m pytest tests/test_body_commands.py tests/test_cpp_api.py -q -m gpu -k "gear_and_mouse_joints_on_running or running_world_creation" 2>&1 | tail -25 | tee gpurun_out/r2N_late_joints.log
  timeout 30 python -m pytest tests/test_body_commands.py -q -m gpu -x -k "world_edits" 2>&1 | tail -12 | tee gpurun_out/r2N_late_joints_regression.log
  exit 0
fi
timeout 70 python -m pytest tests/test_body_commands.py tests/test_cpp_api.py -q -m gpu -k "set_density or running_world_creation" 2>&1 | tail -25 | tee gpurun_out/r2N_density.log
timeout 60 python -m pytest tests/test_body_commands.py -q -m gpu -x -k "world_edits or bad_command or create_and_destroy" 2>&1 | tail -12 | tee gpurun_out/r2N_edits_regression.log
