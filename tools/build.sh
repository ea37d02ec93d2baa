#!/bin/bash
# development aid: build the CUDA library + host port, print register/stack usage and code size
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; import enlsip_jl_b200 as E; E.capi.build(force=True, verbose=True); g.build_hostport(force=True)" > /tmp/enl_build.log 2>&1
echo "build rc=$?"
grep -E "error|Used" /tmp/enl_build.log | sort | uniq -c
# the per-phase probe of the persistent QRCP panel kernel (profiles/r2_qr_panel_persist_phases.txt)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DQP_PROF -I enlsip.jl_b200/csrc tools/qp_probe.cu -o tools/qp_probe.bin
