mkdir -p gpurun_out
for cfg in "HALO=0 DEBUG=3 2CTA=0" "HALO=0 DEBUG=4 2CTA=0" "HALO=0 DEBUG=0 2CTA=0 BN=64" "HALO=0 DEBUG=3 2CTA=0 BN=64"; do
  echo "=== $cfg"; env $(echo $cfg | sed 's/\([A-Z0-9]*\)=/MAPNET_TC_\1=/g') timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7 | cut -c1-100
done
