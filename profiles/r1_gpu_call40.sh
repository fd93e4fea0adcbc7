DWGPU_LIB=$PWD/alt/libdwgpu_timers.so python profiles/phase_times.py B > gpurun_out/r1m_phase_times_B.txt 2> gpurun_out/r1m_phase.err; echo "rc=$?"
tail -n 20 gpurun_out/r1m_phase_times_B.txt
