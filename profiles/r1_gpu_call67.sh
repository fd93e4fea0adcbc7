set -x
timeout 55 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not (fixed_pi_parity or cluster_path_dense or config_b_shape or mailboxes or packed_record or two_live or mixed_shapes or edge_cases or full_size or config_c)" 2>&1 | tail -4
