set -x
timeout 45 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 75 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixed_pi_parity or cluster_path_dense or config_b_shape or mailboxes or packed_record" 2>&1 | tail -4
