#!/bin/bash
# Round-2 first GPU call: environment probe of the B200 box (ffmpeg? NUMA? topology?), the round-1
# test suite as a sanity baseline, and compute-sanitizer memcheck / racecheck over the kernels that use
# shared-memory atomics, union-find and mbarriers.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
{
  echo "== which ffmpeg / ffprobe"; which ffmpeg ffprobe; echo "rc=$?"
  echo "== ldconfig avfilter/avcodec/swscale"; ldconfig -p | grep -i -E "avfilter|avcodec|swscale|avformat" || echo none
  echo "== find libav*"; find / -xdev \( -name "libavfilter*" -o -name "libswscale*" -o -name "ffmpeg*" \) -not -path "/proc/*" 2>/dev/null | head -20
  echo "== python av / imageio_ffmpeg"; python - <<'EOF'
import importlib, shutil
for m in ("av", "imageio_ffmpeg", "ffmpeg", "skvideo", "decord", "PyNvVideoCodec", "nvidia.dali"):
    try:
        importlib.import_module(m); print(m, "importable")
    except Exception as e:
        print(m, "missing:", type(e).__name__)
print("shutil.which ffmpeg:", shutil.which("ffmpeg"), "ffprobe:", shutil.which("ffprobe"))
import cv2
bi = cv2.getBuildInformation()
for ln in bi.splitlines():
    if any(k in ln for k in ("FFMPEG", "avcodec", "avformat", "avutil", "swscale", "Version control", "IPP", "Parallel framework")):
        print(ln.strip())
print("cv2", cv2.__version__, "threads", cv2.getNumThreads())
EOF
  echo "== numactl / lscpu"; which numactl; lscpu | grep -i -E "model name|socket|numa|^CPU\(s\)|thread"
  echo "== nvidia-smi topo"; nvidia-smi topo -m
  echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,pci.bus_id,memory.total --format=csv
  echo "== mem"; free -g | head -2
} > $O/r02_env_probe.txt 2>&1

timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_baseline.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_gpu_baseline.log

CS=/usr/local/cuda/bin/compute-sanitizer
SEL='canny_bit_exact or histograms_bit_exact or dct_coefficients or small_and_ragged or psnr_ssim or gray_all_sizes or resize_bit_exact'
timeout 1500 $CS --tool memcheck --error-exitcode 7 --log-file $O/r02_sanitizer_memcheck.txt \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/r02_sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/r02_sanitizer_memcheck_pytest.log
timeout 1500 $CS --tool memcheck --error-exitcode 7 --log-file $O/r02_sanitizer_memcheck_orb.txt \
    python -m pytest tests/test_orb_gpu.py -m gpu -x -q -k "pyramid_levels or keypoint_sets_match_oracle or batches_device" > $O/r02_sanitizer_memcheck_orb_pytest.log 2>&1
echo "memcheck-orb rc=$?" >> $O/r02_sanitizer_memcheck_orb_pytest.log
SELR='canny_bit_exact or histograms_bit_exact or dct_coefficients'
timeout 1500 $CS --tool racecheck --racecheck-report all --error-exitcode 7 --log-file $O/r02_sanitizer_racecheck.txt \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SELR" > $O/r02_sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> $O/r02_sanitizer_racecheck_pytest.log
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 7 --log-file $O/r02_sanitizer_racecheck_orb.txt \
    python -m pytest tests/test_orb_gpu.py -m gpu -x -q -k "keypoint_sets_match_oracle" > $O/r02_sanitizer_racecheck_orb_pytest.log 2>&1
echo "racecheck-orb rc=$?" >> $O/r02_sanitizer_racecheck_orb_pytest.log
timeout 900 $CS --tool synccheck --error-exitcode 7 --log-file $O/r02_sanitizer_synccheck.txt \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SELR or farneback_flow" > $O/r02_sanitizer_synccheck_pytest.log 2>&1
echo "synccheck rc=$?" >> $O/r02_sanitizer_synccheck_pytest.log
tail -3 $O/r02_pytest_gpu_baseline.log $O/r02_sanitizer_*_pytest.log
for f in $O/r02_sanitizer_*.txt; do echo "== $f"; tail -4 $f; done
