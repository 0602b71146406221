#!/bin/bash
# Round 2, last GPU call: b2Fixture::SetDensity on the device (new parity test, the C++ mirror's program), then the edit
# regression tests that share k_body_commands.
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_body_commands.py tests/test_cpp_api.py -q -m gpu -k "set_density or running_world_creation" 2>&1 | tail -25 | tee gpurun_out/r2N_density.log
timeout 60 python -m pytest tests/test_body_commands.py -q -m gpu -x -k "world_edits or bad_command or create_and_destroy" 2>&1 | tail -12 | tee gpurun_out/r2N_edits_regression.log
