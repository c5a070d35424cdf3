# DRAM / L2 traffic of one full launch of the persistent attack kernel under the switches named on each line
for cfg in "RABI_TC_L2_CARVE_PCT=100" "RABI_TC_L2_CARVE_PCT=200" "RABI_TC_L2_CARVE_PCT=400"; do
  echo "== $cfg"
  env $cfg timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k k_tc -s 1 -c 1 python tools/run_dbg.py lotka_volterra_l2pgd_rkl 10000 200 2>&1 | grep -E "dram__|gpu__time|lts__"
  env $cfg python tools/run_dbg.py lotka_volterra_l2pgd_rkl 10000 200 2>&1 | tail -1
done
