set -x
mkdir -p gpurun_out/r2f
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan3 --launch-skip 4 -c 2 -o gpurun_out/r2f/scan3_fp32 -f python scripts/prof_one.py > gpurun_out/r2f/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r2f/scan3_fp32.ncu-rep > gpurun_out/r2f/scan3_fp32_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan3 --launch-skip 4 -c 2 -o /tmp/scan3_bf16 -f python scripts/prof_one.py bf16 > gpurun_out/r2f/ncu_full_bf16.log 2>&1; echo "ncu full bf16 rc=$?"
python scripts/ncu_summary.py /tmp/scan3_bf16.ncu-rep > gpurun_out/r2f/scan3_bf16_summary.txt 2>&1
python scripts/ncu_top.py /tmp/scan3_bf16.ncu-rep 0 50 > gpurun_out/r2f/scan3_bf16_top_fwd.txt 2>&1
python scripts/ncu_top.py /tmp/scan3_bf16.ncu-rep 1 50 > gpurun_out/r2f/scan3_bf16_top_bwd.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out/r2f
