for cfg in "--m 96 --n 21 --k 4 --norm kl" "--m 96 --n 21 --k 4 --norm fro" "--m 1024 --n 256 --k 4 --norm kl" "--m 1024 --n 256 --k 4 --norm fro"; do
  timeout 120 python tools/tiny_profile.py $cfg 2>&1 | tail -2
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/tiny_kl_96.csv python tools/tiny_profile.py --m 96 --n 21 --k 4 --norm kl --eager-only > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/tiny_kl_1024.csv python tools/tiny_profile.py --m 1024 --n 256 --k 4 --norm kl --eager-only > /dev/null 2>&1
echo done
