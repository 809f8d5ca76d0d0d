import json, sys
for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/variants.jsonl'):
    d=json.loads(l)
    k=d['kernel_ms_per_step']
    print("%-6s P=%-4d fl=%d env=%-48s step %.3f ms b2b %.3f fps %8.0f us/f %6.1f | su %.3f ma %.3f co %.3f ex %.3f | ch %.3f rec %.0f util %s" % (d['workload'],d['poses'],d['flags'],str(d['env']).replace("'",""),d['ms_per_step_flushed'],d['ms_per_step_back_to_back'],d['frames_per_s'],d['us_per_frame_b2b'],k['setup'],k['march'],k['colour'],k['expand'],d['chunks_frac'],d['records_per_frame'],d.get('paint_lane_utilisation')))
