mkdir -p gpurun_out
for cfg in "HALO=1 DEBUG=0" "HALO=1 DEBUG=1" "HALO=1 DEBUG=2" "HALO=0 DEBUG=0 2CTA=0" "HALO=0 DEBUG=1 2CTA=0" "HALO=0 DEBUG=2 2CTA=0" "HALO=0 DEBUG=0 2CTA=1"; do
  echo "=== $cfg"; env $(echo $cfg | sed 's/\([A-Z0-9]*\)=/MAPNET_TC_\1=/g') timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7
done
