#!/bin/bash
# static SASS instruction count per kernel (rough proxy for the per-thread path length)
cuobjdump -sass "${1:-/root/repo/fluid_b200/libfluidb200.so}" | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -n
