#!/bin/bash
# ptxas -v of every kernel (one translation unit at a time) and a SASS digest of the hot kernels.
# Usage: tools/ptxas_report.sh > profiles/rNN_ptxas_v.txt
HERE="$(cd "$(dirname "$0")" && pwd)"
cd "$HERE/../fsilbm3d_b200/csrc"
for f in fluid_kernels.cu ibm_kernels.cu refine_kernels.cu io_kernels.cu fsilbm_api.cu; do
  echo "== $f"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xptxas -v -c -o /tmp/_ptxas_$f.o $f 2>&1 | c++filt | python3 "$HERE/ptxas_parse.py"
done
echo
echo "== SASS digest (cuobjdump -sass of libfsilbm_b200.so): instruction counts per kernel"
cuobjdump -sass ../libfsilbm_b200.so | c++filt | awk '
  /Function :/ { name=$0; sub(/.*Function : /,"",name) }
  /LDG|STG|LDS|STS|RED|ATOMG|LDL|STL|FMA|DFMA|DADD|DMUL|MUFU.RCP64H|BAR/ {
     op=""; for(i=1;i<=NF;i++) if ($i ~ /^(LDG|STG|LDS|STS|RED|ATOMG|LDL|STL|DFMA|DADD|DMUL|BAR)/) { op=$i; break }
     if (op!="") cnt[name" "op]++ }
  END { for (k in cnt) print cnt[k], k }' | grep -E "collide_push_kernel<1, |collide_push_kernel<3, false, false|ibm_loop_kernel<4>|pair_f2s_kernel|pair_extract_kernel" | sort -k2,2 -k1,1nr | awk '{n=$1; $1=""; printf "%6d %s\n", n, $0}'
