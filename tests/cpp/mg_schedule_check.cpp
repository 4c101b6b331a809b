// Host-only check of the V-cycle schedule (petibm_b200/csrc/mg_schedule.h) for every polynomial degree a user can ask
// for: a recording launcher verifies, launch by launch, that no kernel writes a buffer it reads (the kernels read their
// neighbours' values, so in-place updates would race), that every buffer read was written before (or is implicitly
// zero), that the coarse correction is still intact when it is prolonged, and that the returned buffer holds the last
// result of its level.  Exit code 0 = all combinations clean.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "mg_schedule.h"

using namespace b200;

struct Recorder
{
    std::map<const double *, int> version;   // buffer -> number of writes so far (absent: never written)
    std::map<const double *, int> level_of;
    int errors = 0, launches = 0;
    void fail(const char *what, int l)
    {
        std::fprintf(stderr, "schedule error at level %d: %s\n", l, what);
        ++errors;
    }
    void rd(const double *p, int l, const char *name)
    {
        if (!p) return fail(name, l);
        if (!version.count(p)) fail("read of a buffer that was never written", l);
    }
    void wr(double *p, int l)
    {
        if (!p) return fail("null output", l);
        if (level_of.count(p) && level_of[p] != l) fail("write to another level's buffer", l);
        version[p]++;
    }
    void first(int l, const double *b, double *dout, double)
    {
        ++launches;
        rd(b, l, "b");
        if (dout == b) fail("first: output aliases b", l);
        wr(dout, l);
    }
    void step(int l, bool xzero, bool dzero, bool prolong, bool last, const double *b, const double *xin, const double *din,
              const double *ec, double *xout, double *dout, double, double)
    {
        ++launches;
        rd(b, l, "b");
        if (!xzero) rd(xin, l, "xin");
        if (!dzero) rd(din, l, "din");
        if (prolong) rd(ec, l + 1, "ec");
        if (xzero && dzero) fail("step without an iterate", l);
        if (xout == xin || xout == din || xout == b || (prolong && xout == ec)) fail("xout aliases an input", l);
        if (!last && (dout == xin || dout == din || dout == b || dout == xout)) fail("dout aliases", l);
        if (last && dout) fail("last step must not produce d", l);
        wr(xout, l);
        if (!last) wr(dout, l);
    }
    // the recorded tail of the cycle, replayed through the same checks (what k_mg_tail does in one launch)
    std::vector<MgOp> tail_ops;
    double *tail_result = nullptr;
    int tail_launches = 0;
    double *tail()
    {
        const int before = launches;
        for (const MgOp &op : tail_ops)
        {
            if (op.kind == 0) first(op.l, op.b, op.dout, op.c2);
            else if (op.kind == 1) step(op.l, op.xzero, op.dzero, op.prolong, op.last, op.b, op.xin, op.din, op.ec, op.xout, op.dout, op.c1, op.c2);
            else restrict(op.l, op.xzero, op.b, op.xin, op.din, op.xout, op.dout);
        }
        tail_launches += launches - before;
        return tail_result;
    }
    void restrict(int l, bool xzero, const double *b, const double *xin, const double *din, double *xsum, double *bc)
    {
        ++launches;
        rd(b, l, "b");
        if (!xzero) rd(xin, l, "xin");
        rd(din, l, "din");
        if (xsum == xin || xsum == din || xsum == b) fail("xsum aliases an input", l);
        wr(xsum, l);
        wr(bc, l + 1);
    }
};

int main()
{
    int bad = 0, combos = 0;
    for (int nl = 1; nl <= 6; ++nl)
        for (int sm = 1; sm <= 6; ++sm)
            for (int co = 1; co <= 6; ++co)
            {
                std::vector<std::vector<double>> store((size_t)nl * 5 + 1, std::vector<double>(1));
                double *work[32][4];
                double *rhs[32];
                Recorder R;
                for (int l = 0; l < nl; ++l)
                {
                    for (int q = 0; q < 4; ++q)
                    {
                        work[l][q] = store[(size_t)l * 5 + q].data();
                        R.level_of[work[l][q]] = l;
                    }
                    rhs[l] = l > 0 ? store[(size_t)l * 5 + 4].data() : nullptr;
                    if (rhs[l]) R.level_of[rhs[l]] = l;
                }
                double *b0 = store.back().data();
                R.version[b0] = 1;
                MgParams prm;
                prm.smooth_its = sm;
                prm.coarse_its = co;
                double *z = mg_cycle(0, nl, b0, work, rhs, prm, R);
                bool ok = z && R.version.count(z) && R.errors == 0;
                bool mine = false;
                for (int q = 0; q < 4; ++q) mine = mine || z == work[0][q];
                ok = ok && mine;
                // launches: coarsest 1 + (co-1); every other level first + (sm-1) + restrict + sm
                const int want = (nl - 1) * (2 * sm + 1) + co;
                if (R.launches != want)
                {
                    std::fprintf(stderr, "nl %d sm %d co %d: %d launches, expected %d\n", nl, sm, co, R.launches, want);
                    ok = false;
                }
                // with the coarse levels recorded and replayed as a tail (every possible tail level): same launches in the
                // same order, same result buffer
                for (int tl = 1; tl < nl && ok; ++tl)
                {
                    MgProgramRecorder rec;
                    MgParams p0 = prm;
                    p0.tail_level = -1;
                    double *tres = mg_cycle(tl, nl, rhs[tl], work, rhs, p0, rec);
                    Recorder RT;
                    RT.level_of = R.level_of;
                    RT.version[b0] = 1;
                    RT.tail_ops = rec.ops;
                    RT.tail_result = tres;
                    MgParams p1 = prm;
                    p1.tail_level = tl;
                    double *zt = mg_cycle(0, nl, b0, work, rhs, p1, RT);
                    ok = ok && zt == z && RT.errors == 0 && RT.launches == want && (int)rec.ops.size() == RT.tail_launches;
                }
                // the schedule is the same in every cycle: a second cycle returns the same buffer
                Recorder R2 = R;
                double *z2 = mg_cycle(0, nl, b0, work, rhs, prm, R2);
                ok = ok && z2 == z && R2.errors == 0;
                ++combos;
                if (!ok)
                {
                    std::fprintf(stderr, "FAILED: levels %d smooth %d coarse %d\n", nl, sm, co);
                    ++bad;
                }
            }
    std::printf("mg schedule: %d combinations, %d bad\n", combos, bad);
    return bad ? 1 : 0;
}
