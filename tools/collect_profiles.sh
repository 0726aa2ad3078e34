# Turns the gpurun_out/ files of tools/profile_r2.sh (+ the BUSY-kernel capture) into the tracked artifacts under profiles/.
# usage: bash tools/collect_profiles.sh [prefix]   (default r2b); run from the repository root after the library was built
P=${1:-r2b}
set -e
for k in dct pack spu xa; do python tools/ncu_summary.py gpurun_out/${P}_$k.ncu-rep > profiles/${P}_ncu_$k.txt; done
[ -f gpurun_out/${P}_busy.ncu-rep ] && python tools/ncu_summary.py gpurun_out/${P}_busy.ncu-rep > profiles/${P}_ncu_pack_busy_hard.txt
for f in bench.json bench_reference.json launches.csv workloads.txt; do cp gpurun_out/${P}_$f profiles/${P}_$f; done
L=psxavenc_b200/libpsxav_b200.so
line() { grep -n -e "$1" psxavenc_b200/csrc/bs_encode.cu | head -1 | cut -d: -f1; }
a=$(line "^__device__ __forceinline__ uint32_t warp_sum"); b=$(line "^template <bool UPPER>"); c=$(line "^// The first pass (q = 1)")
d=$(line "^// The same for a dense group"); e=$(line "^// Emit, convergent part for a dense group"); f=$(line "^// Emit, convergent part: parks")
g=$(line "^// Appends MSB-first codes"); h=$(line "^// v3 DC prediction chain"); i=$(line "^// ---- census"); j=$(line "^// STR mode: where frame f")
k=$(line "---- (1) first-fit quant scale search"); l=$(line "---- (2) exclusive scan"); m=$(line "---- (3) emit ---"); n=$(line "---- (4) header, results")
o=$(line "^// ---- launchers"); k1=$(line "^bs_dct_kernel(const uint8_t"); k2=$(line "int v\[64\];"); k3=$(line "fdct8x8<VARIANT>(v);")
k4=$(line "uint32_t sign_lo = 0"); k5=$(line "const int count = (int)(tail"); k6=$(line "^// ---- kernel 2")
{
echo "Phase attribution (tools/ncu_lines.py: SASS page of the ncu report joined with the cubin's line table), strv typical content"
echo "(noise 3, q = 2), 4096 frames per launch, kernels at the end of round 2:"
python tools/ncu_lines.py gpurun_out/${P}_pack.ncu-rep $L 'bs_pack_kernel<false, true, false, false, 320, 4, 1>' $a:$((b-1)):helpers-warp_sum/imad \
  $b:$((c-1)):pricing-lists-generic $c:$((d-1)):pricing-lists-q1 $d:$((e-1)):pricing-dense $e:$((f-1)):emit-stage_dense $f:$((g-1)):emit-stage_rows \
  $g:$((h-1)):emit-BitWriter $j:$((k-1)):setup $k:$((l-1)):search-loop-body $l:$((m-1)):scan $m:$((n-1)):emit-loop $n:$((o-1)):copy-out
python tools/ncu_lines.py gpurun_out/${P}_dct.ncu-rep $L 'bs_dct_kernel<1>' $k1:$((k2-1)):setup+zero-lists $k2:$((k3-1)):gather $k3:$((k4-1)):fdct-call \
  $k4:$((k5-1)):sign+y+list-append $k5:$((k6-1)):rows-out
echo; echo "fdct.cuh / intrinsics lines (the transform itself) make up the rest of bs_dct_kernel."
if [ -f gpurun_out/${P}_busy.ncu-rep ]; then
echo; echo "BUSY instantiation of the pack kernel on noise-6 content (q = 8), tools/wave_probe.py 6 4096 (profiles/${P}_ncu_pack_busy_hard.txt):"
python tools/ncu_lines.py gpurun_out/${P}_busy.ncu-rep $L 'bs_pack_kernel<false, true, false, true, 320, 4, 1>' $a:$((b-1)):helpers-warp_sum/imad \
  $d:$((e-1)):pricing-dense $e:$((f-1)):emit-stage_dense $g:$((h-1)):emit-BitWriter $i:$((j-1)):census $j:$((k-1)):setup $k:$((l-1)):search-loop-body \
  $m:$((n-1)):emit-loop $n:$((o-1)):copy-out
fi
} > profiles/${P}_phases.txt 2>&1
mkdir -p profiles/${P}_sass
for spec in "bs_dct_kernelILi1E:bs_dct_kernel_sse2" "bs_pack_kernelILb0ELb1ELb0ELb0ELi320ELi4E:bs_pack_kernel_v2_320x4" \
            "bs_pack_kernelILb0ELb1ELb0ELb1ELi320ELi4E:bs_pack_kernel_v2_busy_320x4" \
            "bs_pack_kernelILb0ELb1ELb0ELb0ELi640ELi1ELi4E:bs_pack_kernel_v2_cluster4" "adpcm_spu_kernelE:adpcm_spu_kernel" \
            "adpcm_spu_small_kernel:adpcm_spu_small_kernel"; do
  cuobjdump -sass $L | awk -v pat="${spec%%:*}" '/Function : /{p=($0 ~ pat)} p' > profiles/${P}_sass/${spec##*:}.sass
done
python - <<PY
import json, re
def traffic(path):
    for l in open(path):
        m = re.search(r"dram traffic per launch \(bytes\): (\d+)", l)
        if m: return int(m.group(1))
t = json.load(open("profiles/traffic.json"))
t["_source"] = "profiles/${P}_ncu_dct.txt, ${P}_ncu_pack.txt, ${P}_ncu_spu.txt (ncu --set full, 4096 frames per launch, fdct sse2, strv noise_bits 3; vagi x 1024): dram__bytes_read.sum + dram__bytes_write.sum per launch"
t["bs_dct_kernel"], t["bs_pack_kernel"], t["adpcm_spu_kernel"] = traffic("profiles/${P}_ncu_dct.txt"), traffic("profiles/${P}_ncu_pack.txt"), traffic("profiles/${P}_ncu_spu.txt")
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
PY
tail -30 profiles/${P}_phases.txt
