set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
timeout 120 tools/qp_probe.bin 4097 3587 > gpurun_out/qp_probe_4097x3587.txt 2>&1
timeout 60 tools/qp_probe.bin 4096 511 > gpurun_out/qp_probe_4096x511.txt 2>&1
timeout 200 python tools/vecop_time.py > gpurun_out/vecop_time.txt 2>&1
ENLSIP_SMALL_PROF=1 timeout 300 python tools/small_one.py c5 4096 7 > gpurun_out/c5_small_prof.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qrcp_launches.csv python tools/small_one.py qrcp 4097 3587 > gpurun_out/qrcp_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_launches.csv python tools/small_one.py c5 4096 2 > gpurun_out/c5_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_panel_persist -s 3 -c 1 -o gpurun_out/qr_panel_persist python tools/small_one.py qrcp 4097 3587 > gpurun_out/qrpp_ncu.log 2>&1
ls -la gpurun_out | head -30
