#!/bin/bash
# usage: sass_count.sh <object> <mangled-name-substring>   -> static SASS instruction count and opcode histogram of one kernel
obj=$1; pat=$2
cuobjdump -sass "$obj" | awk -v pat="$pat" '/Function : /{on = index($0, pat) > 0} on' > /tmp/_k.sass
echo "instructions: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/_k.sass)"
grep -o '^\s*/\*[0-9a-f]\{4\}\*/\s*\(@!\?U\?P[0-9T]\s\)\?\s*[A-Z0-9_.]*' /tmp/_k.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-14} | tr '\n' ' '; echo
