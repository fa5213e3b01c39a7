import json,sys,glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); r=d['roofline']
        print(f"{f.split('/')[-1]:28s} rays/s {d['value']/1e6:6.2f}M step {d['ms_per_step']:7.2f}ms gather {r['avg_launch_ms']:7.2f} march {r.get('march_kernels_ms',0):6.2f} frac {r['frac']:.3f} cand {r['candidates_per_lookup']:.1f} found {r['photons_found_per_lookup']:.2f} e2e {d['e2e']['value']/1e6:.2f}M chk {d['checksum_L']}")
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-800:])
