nproc; lscpu | grep -E "Model name|Thread|Core|Socket|L2|L3"
for t in 1 2 4 8 16; do alt/team_bench $t; done
