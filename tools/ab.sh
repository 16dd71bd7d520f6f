#!/bin/bash
# kernel-variant A/B: run the device-resident quick bench once per "library[:impl]" given on the command line
# (impl = warp | cta, the DCN_FUSED_IMPL switch of the library)
for spec in "$@"; do
  lib=${spec%%:*}; impl=${spec#*:}; [ "$impl" = "$spec" ] && impl=warp
  echo "== $lib impl=$impl"
  DCN_FUSED_IMPL=$impl DCN_LIB=$PWD/deacon_server_b200/$lib timeout 600 python tools/quick_bench.py --genome-mbp 20 --pad-keys-m 360 --pairs-m 5 --check 20000 --no-e2e 2>&1 | grep -E "device-resident|parity|Error|error|PARITY|probe rate|ceiling"
done
