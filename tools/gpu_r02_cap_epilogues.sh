mkdir -p gpurun_out
for pair in "dgelu:5" "fc1dual:6"; do
  name=${pair%%:*}; epi=${pair##*:}
  timeout 120 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k "regex:gemm_tcgen05_kernel<\(int\)256, \(int\)${epi}, \(int\)2" -c 1 -f -o gpurun_out/prof_$name python tools/train_bench.py --batch 4096 --skip-kernels --profile-step > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_table.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_${name}_table.md 2>>gpurun_out/ncu_$name.log
  tail -1 gpurun_out/ncu_${name}_table.md
done
rm -f gpurun_out/*.ncu-rep
