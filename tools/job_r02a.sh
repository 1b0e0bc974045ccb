set -x
cd $GRAFT_REPO_ROOT
python tools/enc_profile.py --cx 0 5 8 10 --sig music --n 65536 > gpurun_out/r02a_enc_profile.txt 2>&1
python tools/enc_profile.py --cx 10 --sig harmonic --n 65536 >> gpurun_out/r02a_enc_profile.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_enc_launches.csv python tools/enc_profile.py --cx 10 --sig music --n 65536 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02a_enc_frame_rdo python tools/enc_profile.py --cx 10 --sig music --n 32768 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_analysis -s 4 -c 1 -f -o gpurun_out/r02a_enc_analysis python tools/enc_profile.py --cx 10 --sig music --n 32768 --reps 1 > /dev/null 2>&1
cat gpurun_out/r02a_enc_profile.txt
