# ncu --set full of the HBM-bound convolutions of the last vocoder stages at B = 64:
# tc_conv_kernel launch indices (profiles/r2_launches.csv): 90/91 = C 24 k 3 (first conv, second conv + residual),
# 102/103 = C 24 k 11, 77/78 = C 48 k 7.
P="python tools/profile_step.py --precision fp16 --batch 64"
F="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
T=${1:-r2b}
$F -k regex:tc_conv_kernel -s 90 -c 2 -o gpurun_out/${T}_prof_c24k3 $P > /dev/null 2>&1
$F -k regex:tc_conv_kernel -s 102 -c 2 -o gpurun_out/${T}_prof_c24k11 $P > /dev/null 2>&1
$F -k regex:tc_conv_kernel -s 77 -c 2 -o gpurun_out/${T}_prof_c48k7 $P > /dev/null 2>&1
ls -la gpurun_out/${T}_prof_c*
