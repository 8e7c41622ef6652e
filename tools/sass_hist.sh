#!/bin/bash
# SASS opcode histogram of one kernel of the built library: tools/sass_hist.sh <kernel name substring>
cuobjdump -sass airdos_b200/lib/libairdos_b200.so | awk -v k="$1" '/Function : /{f=index($0,k)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn | head -${2:-36} | paste - - - -
