#!/bin/bash
# Per-kernel SASS mnemonic counts of the built library (what proves TMA stores, setmaxnreg warp specialisation, packed f32x2
# arithmetic, redux votes).  usage: tools/sass_evidence.sh > profiles/r2_sass_evidence.txt
LIB=spectrogram-midi_b200/libaegis_b200.so
echo "# cuobjdump -sass $LIB  (nvcc $(nvcc --version | grep release | sed 's/.*release //'), sm_100a)"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
{ for (i=1;i<=NF;i++) { op=$i; sub(/\..*/,"",op);
    if (op=="UTMASTG"||op=="UTMALDG"||op=="USETMAXREG"||op=="FFMA2"||op=="FADD2"||op=="FMUL2"||op=="LDGSTS"||op=="CREDUX"||op=="REDUX"||op=="DADD"||op=="DSETP"||op=="ATOMS"||op=="BAR") c[fn" "op]++ } }
END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { if (last!="") print ""; printf "%s:", $1; last=$1 } printf " %s=%s", $2, $3 } END { print "" }' | c++filt 2>/dev/null | sed 's/(aegis_[a-z_]*params[^)]*)//'
