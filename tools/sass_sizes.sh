#!/bin/bash
# SASS instruction count per kernel of the built library (16 bytes each): the instruction-cache footprint
cuobjdump -sass ${1:-/root/repo/collision_detect_gjk_epa_b200/libgjkepa_b200.so} | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -n
