#!/bin/bash
# SASS instruction counts per function of a built library: sass_sizes.sh <lib.so>
cuobjdump -sass "$1" | awk '/Function : /{fn=$3} /^[ \t]+\/\*[0-9a-f]+\*\/ /{c[fn]++} END{for(f in c) print c[f], f}' | sort -rn | c++filt | cut -c1-110 | head -${2:-14}
