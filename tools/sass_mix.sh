#!/bin/bash
# SASS instruction mix of one kernel: tools/sass_mix.sh <object.o> <kernel name substring>
cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f=$3} f ~ pat && /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/{ op=$2; if (op ~ /^@/) op=$3; sub(/[.;].*/,"",op); print op }' | sort | uniq -c | sort -rn
