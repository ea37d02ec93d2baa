#!/bin/bash
# development aid: build the CUDA library + host port, print register/stack usage and code size
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; import enlsip_jl_b200 as E; E.capi.build(force=True, verbose=True); g.build_hostport(force=True)" > /tmp/enl_build.log 2>&1
echo "build rc=$?"
grep -E "error|Used" /tmp/enl_build.log | sort | uniq -c
