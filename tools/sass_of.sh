#!/bin/bash
# usage: tools/sass_of.sh <kernel-name-substring> [lib]  -> plain SASS listing (addr instr) on stdout
LIB=${2:-rawspeed_b200/librawspeed_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v k="$1" '
/Function :/ { on = (index($0, k) > 0) }
on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ { sub(/^[ \t]+\/\*/, ""); sub(/\*\/[ \t]+/, " "); sub(/[ \t]*\/\*.*$/, ""); print }'
