#!/bin/bash
# repeats the tests whose outcome depends on the order of fp32 atomics and prints their measured quantities (-s output)
for i in $(seq 1 ${1:-8}); do
  python -m pytest -q -s -x tests/test_parity_full_gpu.py::test_loss_trajectory_fresh_batches_with_dropout tests/test_modules_gpu.py::test_loss_trajectory_100_steps tests/test_graph_ddp_gpu.py::test_graphed_training_step_matches_eager tests/test_graph_ddp_gpu.py::test_fused_optimizer_state_dict_roundtrip tests/test_backward_gpu.py::test_planar_activation_mirror_gives_the_same_gradients tests/test_backward_gpu.py::test_block_fused_backward_matches_the_per_layer_backward "tests/test_backward_gpu.py::test_whole_gradient_vs_fp32_oracle_and_autocast_envelope" tests/test_ops_gpu.py::test_dense_block_backward_is_deterministic_enough 2>&1 | grep -E "oracle loss drop|max \|ours|FAILED|passed|failed|rel|cos|differ" | cut -c1-220
done
