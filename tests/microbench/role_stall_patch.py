"""Diagnostic patch of csrc/ (apply, build, measure, undo -- never commit the patched sources): the producer /
MMA / epilogue roles count the cycles they wait on each mbarrier, and B200_DBG bits skip the TMA loads (1), the
MMAs (2) or the epilogue body (4).

    python tests/microbench/role_stall_patch.py [--csrc DIR]        apply (keeps *.orig copies)
    python tests/microbench/role_stall_patch.py [--csrc DIR] undo   restore

tests/test_abi.py applies it to a scratch copy of csrc/ so that it cannot silently fall behind the sources."""
import os, shutil, sys
d = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'bilinear_b200', 'csrc') + os.sep
if '--csrc' in sys.argv:
    i = sys.argv.index('--csrc')
    d = os.path.join(sys.argv[i + 1], '')
    del sys.argv[i:i + 2]
if sys.argv[1:] == ['undo']:
    for f in ['mega.cu','api.cu','jobs.h']: shutil.copy(d+f+'.orig', d+f)
    sys.exit(0)
for f in ['mega.cu','api.cu','jobs.h']: shutil.copy(d+f, d+f+'.orig')
def rep(s, a, b, cnt=1):
    assert a in s, a[:60]
    return s.replace(a, b) if cnt == 0 else s.replace(a, b, cnt)
j=open(d+'jobs.h').read()
j=rep(j,"  uint32_t bar_base;        // value of *barrier when this launch starts (the host keeps count)\n","  uint32_t bar_base;        // value of *barrier when this launch starts (the host keeps count)\n  int32_t dbg;\n")
i=j.index('struct Dyn {'); e=j.index('};',i)
j=j[:e]+"  int32_t dbg_skip_epi;\n"+j[e:]
open(d+'jobs.h','w').write(j)
a=open(d+'api.cu').read()
a=rep(a,"  p.dyn = dyn;\n","  p.dyn = dyn;\n  p.dbg = getenv(\"B200_DBG\") ? atoi(getenv(\"B200_DBG\")) : 0;\n  p.dyn.dbg_skip_epi = p.dbg & 4;\n")
a=rep(a,"    for (int k = 1; k < 8; ++k) ns_out[8 * i + k] = t[8 * i + k] >= t[8 * i] ? int64_t(t[8 * i + k] - t[8 * i]) : -1;","    for (int k = 1; k < 8; ++k) ns_out[8 * i + k] = int64_t(t[8 * i + k]);")
open(d+'api.cu','w').write(a)
s=open(d+'mega.cu').read()
s=rep(s,"SmemCtl* ctl, Ring& pipe, int rank, uint32_t elected) {\n  // The operand layouts are compile-time here","SmemCtl* ctl, Ring& pipe, int rank, uint32_t elected, long long& stall) {\n  // The operand layouts are compile-time here")
s=rep(s,"SmemCtl* ctl, Ring& pipe, int rank, uint32_t elected) {\n  const int a_mn = tr.job->a_mn, b_mn = tr.job->b_mn;","SmemCtl* ctl, Ring& pipe, int rank, uint32_t elected, long long& stall) {\n  const int a_mn = tr.job->a_mn, b_mn = tr.job->b_mn;")
s=rep(s,"(p, phs, tr, smem, ctl, pipe, rank, elected);","(p, phs, tr, smem, ctl, pipe, rank, elected, stall);",0)
s=rep(s,'    mbar_wait(&ctl->empty[pipe.stage], pipe.parity(), 100);\n    if (elected) {\n      const uint32_t full =','    { long long t0 = clock64(); mbar_wait(&ctl->empty[pipe.stage], pipe.parity(), 100); stall += clock64() - t0; }\n    if (p.dbg & 1) { if (elected && (CG == 1 || rank == 0)) mbar_arrive(&ctl->full[pipe.stage]); pipe.advance(phs.num_stages); continue; }\n    if (elected) {\n      const uint32_t full =')
s=rep(s,'                                         Pipe& acc, uint32_t tmem_base, uint32_t elected) {','                                         Pipe& acc, uint32_t tmem_base, uint32_t elected, long long& stall_full, long long& stall_acc, long long& t_issue, int dbg) {')
s=rep(s,'  mbar_wait(&ctl->tmem_empty[acc.stage], acc.phase ^ 1u, 200);\n','  { long long t0 = clock64(); mbar_wait(&ctl->tmem_empty[acc.stage], acc.phase ^ 1u, 200); stall_acc += clock64() - t0; }\n')
s=rep(s,'      mbar_wait(&ctl->full[pipe.stage], pipe.parity(), 201);\n','      long long t0 = clock64(); mbar_wait(&ctl->full[pipe.stage], pipe.parity(), 201); long long t1 = clock64(); stall_full += t1 - t0;\n', 0)
s=rep(s,"      if (elected) {\n        umma_bf16<CG>(tmem_d, da, db, idesc, accumulate);","      if (elected && !(dbg & 2)) {\n        umma_bf16<CG>(tmem_d, da, db, idesc, accumulate);")
s=rep(s,"        umma_commit<CG>(&ctl->empty[pipe.stage]);  // frees the smem stage once these MMAs have read it\n      }\n","        umma_commit<CG>(&ctl->empty[pipe.stage]);  // frees the smem stage once these MMAs have read it\n      }\n      if (elected && (dbg & 2)) umma_commit<CG>(&ctl->empty[pipe.stage]);\n      t_issue += clock64() - t1;\n")
s=rep(s,'uint32_t tmem_base, int warp, int lane, int rank) {\n  const GemmJob* job = tr.job;\n  mbar_wait(&ctl->tmem_full[acc.stage], acc.phase, 300);','uint32_t tmem_base, int warp, int lane, int rank, long long& stall) {\n  const GemmJob* job = tr.job;\n  { long long t0 = clock64(); mbar_wait(&ctl->tmem_full[acc.stage], acc.phase, 300); stall += clock64() - t0; }')
s=rep(s,"  switch (job->epi) {\n    case EPI_STORE_F32: epi_store_f32(job, taddr, f, n0); break;","  if (!(d.dbg_skip_epi)) switch (job->epi) {\n    case EPI_STORE_F32: epi_store_f32(job, taddr, f, n0); break;")
s=rep(s,"        if (stamp && lane == 0) p.phase_ns[8 * ph + 1] = global_timer_ns();\n","        long long stall = 0; const long long c0 = clock64();\n")
s=rep(s,'producer_tile<CG>(p, phase, tr, smem, ctl, prod, rank, elected);','producer_tile<CG>(p, phase, tr, smem, ctl, prod, rank, elected, stall);')
s=rep(s,"        if (stamp && lane == 0) p.phase_ns[8 * ph + 2] = global_timer_ns();","        if (stamp && lane == 0) { p.phase_ns[8 * ph + 1] = stall; p.phase_ns[8 * ph + 2] = clock64() - c0; }")
s=rep(s,"          Ring cons{0, ctl->cons_bits};","          long long sfull = 0, sacc = 0, tiss = 0;\n          Ring cons{0, ctl->cons_bits};")
s=rep(s,'mma_tile<CG>(phase, tr, smem, ctl, cons, acc_mma, tmem_base, elected);','mma_tile<CG>(phase, tr, smem, ctl, cons, acc_mma, tmem_base, elected, sfull, sacc, tiss, p.dbg);')
s=rep(s,"            if (stamp) p.phase_ns[8 * ph + 3] = global_timer_ns();","            if (stamp) { p.phase_ns[8 * ph + 3] = sfull; p.phase_ns[8 * ph + 5] = tiss; if (ph > 0) p.phase_ns[8 * ph + 7] = sacc; }")
s=rep(s,"""        bool first = true;
        for (TileIter it(p, ctl, CG); it.valid(); it.next()) {
          TileRef tr = find_tile(blob, it.tile());
          if (stamp && first && warp == 4 && lane == 0) {
            mbar_wait(&ctl->tmem_full[acc_epi.stage], acc_epi.phase, 301);
            p.phase_ns[8 * ph + 4] = global_timer_ns();
          }
          first = false;
          epilogue_tile<CG>(tr, p.dyn, ctl, acc_epi, tmem_base, warp, lane, rank);
        }
        if (stamp && warp == 4 && lane == 0) p.phase_ns[8 * ph + 5] = global_timer_ns();""","""        long long stall = 0;
        for (TileIter it(p, ctl, CG); it.valid(); it.next()) {
          TileRef tr = find_tile(blob, it.tile());
          epilogue_tile<CG>(tr, p.dyn, ctl, acc_epi, tmem_base, warp, lane, rank, stall);
        }
        if (stamp && warp == 4 && lane == 0) p.phase_ns[8 * ph + 4] = stall;""")
open(d+'mega.cu','w').write(s)
print("patched")
