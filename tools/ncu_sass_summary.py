import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_pg_source.csv')))
H = rows[1]
isrc, isamp, iexec = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
data = []
for r in rows[2:]:
    if len(r) < len(H): continue
    try: data.append((int(r[isamp] or 0), int(r[iexec] or 0), r[isrc]))
    except ValueError: pass
# the capture holds two launches (full waves, tail wave) and the page repeats the SASS per launch: keep the first copy
for p in range(1000, len(data)):
    if all(data[p + k][2] == data[k][2] for k in range(50)):
        data = data[:p]
        break
print('ncu --set full --import-source on, source page (SASS) of tc_rollout_kernel<PathTracking,BWD>, full-wave launch (444 tiles, grid 148), B = 65,536, n = 25, final round-2 kernel')
print('SASS instructions %d, warp-level instructions executed %d, stall samples %d' % (len(data), sum(e for _, e, _ in data), sum(s for s, _, _ in data)))
base = 444*26*16*4
print()
print('Epilogue block loops (executed once per tile, step, epilogue warp and 64-feature block = %d times each; 16 elements per thread per block):' % base)
segs = []; cur = None
for i, (s, e, src) in enumerate(data):
    if e == base:
        if cur is None: cur = [i, i, 0]
        cur[1] = i; cur[2] += s
    else:
        if cur and i - cur[1] > 150: segs.append(cur); cur = None
if cur: segs.append(cur)
tot = 0
big = [sg for sg in segs if sum(1 for i in range(sg[0], sg[1]+1) if data[i][1] == base) >= 50]
names = ['E1   (z1 chunk -> ELU -> fp16 pair image of h1)', 'E2   (z2 -> ELU -> head dot products, h2 pair to the h2 store)', 'Ed2  (h2 image -> delta2 pair image, in place)', 'Ed1  (g_h1, z1 chunk -> delta1 pair image)']
for k, (a, b, smp) in enumerate(big):
    idx = [i for i in range(a, b+1) if data[i][1] == base]
    n = len(idx); tot += n
    c = lambda key: sum(1 for i in idx if key in data[i][2])
    print('  SASS %5d-%5d  %-66s %4d instr/block = %4.1f per element; MUFU %2d LDS %2d STS %d STG %d LDTM %d; stall samples %d' % (a, b, names[k] if k < 4 else '?', n, n/16, c('MUFU'), c('LDS'), c('STS'), c('STG'), c('LDTM'), smp))
print('  => %d warp instructions per (tile, step, warp) in the four epilogues; 4 epilogue warps per scheduler: %.1f K issue cycles per forward + backward step of ~38 K cycles' % (4*tot, 16*tot/1e3))
n_mma = sum(1 for _, _, s in data if 'UTCHMMA' in s); n_loop = sum(1 for _, _, s in data if 'BRA.U.ANY' in s)
print()
print("UTCHMMA instructions in the kernel: %d; BRA.U.ANY (per-instruction ELECT loops; what is left of them sits in the elected epilogue thread's bulk stores): %d" % (n_mma, n_loop))
print()
print('Top 30 instructions by stall samples (NANOSLEEP.SYNCS = inside an mbarrier wait):')
ts = sum(s for s,_,_ in data)
for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:30]:
    print('  SASS %5d  samples %6d (%4.1f%%)  executed %9d   %s' % (i, d[0], 100*d[0]/ts, d[1], d[2][:100]))
