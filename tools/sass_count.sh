#!/bin/bash
# static SASS statistics of one kernel: total instructions and the MOV / SEL family (no GPU needed)
# usage: tools/sass_count.sh <mangled-name-substring>
so=${2:-vae_mdl_b200/libvaemdl_b200.so}
cuobjdump -sass "$so" 2>/dev/null | awk -v pat="$1" '/Function : /{f=($0 ~ pat)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" \
 | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' \
 | awk '{n++; split($1,a,"."); op=a[1]; if($1 ~ /^IMAD\.MOV/) op="IMAD.MOV"; c[op]++} END{printf "total %d | MOV %d IMAD.MOV %d SEL %d FSEL %d FMUL2 %d FFMA2 %d FADD2 %d BSSY %d LOP3 %d FSETP %d\n", n, c["MOV"], c["IMAD.MOV"], c["SEL"], c["FSEL"], c["FMUL2"], c["FFMA2"], c["FADD2"], c["BSSY"], c["LOP3"], c["FSETP"]}'
