#!/bin/bash
# kernel-variant A/B: run the device-resident quick bench once per library given on the command line
for lib in "$@"; do
  echo "== $lib"
  DCN_LIB=$PWD/deacon_server_b200/$lib python tools/quick_bench.py --genome-mbp 20 --pad-keys-m 360 --pairs-m 5 --check 20000 --no-e2e 2>&1 | grep -E "device-resident|parity|Error|error"
done
