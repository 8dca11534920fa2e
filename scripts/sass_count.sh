#!/bin/bash
# usage: scripts/sass_count.sh "<extra nvcc flags>"  -- instruction mix of one unit of the fused kernel's epilogue (kGroups = 1)
set -e
cd /root/repo/poserisk_release_b200/csrc
mkdir -p /tmp/sass
nvcc $1 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v -c prk_fused.cu -o /tmp/sass/v.o 2>&1 | grep -A2 "kernelILi1" | tail -2
cuobjdump -sass /tmp/sass/v.o | awk '/Function : .*kernelILi1/{f=1} /Function : .*identity/{f=0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/sass/v.txt
first=$(grep -n "LDTM" /tmp/sass/v.txt | head -1 | cut -d: -f1)
last=$(grep -n "STG" /tmp/sass/v.txt | tail -1 | cut -d: -f1)
echo "unit body: lines $first..$last = $((last-first+1)) instructions"
awk -v a=$first -v b=$last 'NR>=a && NR<=b' /tmp/sass/v.txt | sed -E 's/^[0-9a-f]{4} (@!?U?P[0-9T] )?//' | awk '{print $1}' | sort | uniq -c | sort -rn | head -${2:-14}
