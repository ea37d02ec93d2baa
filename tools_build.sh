#!/bin/bash
# development aid: build the CUDA library + host port, print register/stack usage and code size
set -e
cd "$(dirname "$0")"
python -c "import __graft_entry__ as g; import enlsip_jl_b200 as E; E.capi.build(force=True, verbose=True); g.build_hostport(force=True)" 2>&1 | grep -E "error|Used|stack frame" | head -6
rm -rf /tmp/cub && mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all "$OLDPWD/enlsip.jl_b200/lib/libenlsip_b200.so" >/dev/null 2>&1; readelf -SW *.cubin 2>/dev/null | grep "\.text\." | awk '{print "text bytes 0x"$6, substr($2,1,100)}')
