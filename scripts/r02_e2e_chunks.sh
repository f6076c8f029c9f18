# end-to-end (host-buffer) path: images per full chunk
for c in 2 4 8; do
  PWC_HOST_CHUNK=$c timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunk', $c, round(e['value'],1), round(e['ms_per_step'],3), round(e['host_link_ceiling']['bare_copy_ms'],3), round(e['frac_of_host_link_ceiling'],3))"
done
