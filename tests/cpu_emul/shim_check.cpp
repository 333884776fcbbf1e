// CPU check of the AMReX stand-in's bulk routines (TEST INFRASTRUCTURE): FillBoundary,
// Copy and the min / max / sum reductions against naive per-cell loops written from
// their definition (AMReX FillBoundary semantics: ghost cells take the periodic image,
// ghosts outside a non-periodic domain stay untouched).
#include <cstdio>
#include <random>

#include <AMReX.H>

using namespace amrex;

static int check(int nx, int ny, int nz, int ngrow, int ncomp, int px, int py, int pz, unsigned seed) {
    Box dom(IntVect(2, -1, 0), IntVect(2 + nx - 1, -1 + ny - 1, nz - 1));      // non-zero lower corner on purpose
    BoxArray ba(dom);
    DistributionMapping dm(ba);
    iMultiFab a(ba, dm, ncomp, ngrow), ref(ba, dm, ncomp, ngrow);
    std::mt19937 rng(seed);
    const Box g = grow(dom, ngrow);
    for (int c = 0; c < ncomp; ++c)
        for (int k = g.smallEnd(2); k <= g.bigEnd(2); ++k)
            for (int j = g.smallEnd(1); j <= g.bigEnd(1); ++j)
                for (int i = g.smallEnd(0); i <= g.bigEnd(0); ++i) a(i, j, k, c) = ref(i, j, k, c) = (int)(rng() % 1000) - 500;
    Periodicity per;
    per.p = {px, py, pz};
    a.FillBoundary(per);
    int bad = 0;
    const int lo[3] = {dom.smallEnd(0), dom.smallEnd(1), dom.smallEnd(2)}, n[3] = {nx, ny, nz};
    for (int c = 0; c < ncomp; ++c)
        for (int k = g.smallEnd(2); k <= g.bigEnd(2); ++k)
            for (int j = g.smallEnd(1); j <= g.bigEnd(1); ++j)
                for (int i = g.smallEnd(0); i <= g.bigEnd(0); ++i) {
                    int s[3] = {i, j, k};
                    bool ghost = false, ok = true;
                    for (int d = 0; d < 3; ++d)
                        if (s[d] < lo[d] || s[d] >= lo[d] + n[d]) {
                            ghost = true;
                            if (!per.isPeriodic(d)) ok = false;
                            else s[d] = lo[d] + (((s[d] - lo[d]) % n[d]) + n[d]) % n[d];
                        }
                    const int expect = (ghost && ok) ? ref(s[0], s[1], s[2], c) : ref(i, j, k, c);
                    if (a(i, j, k, c) != expect) ++bad;
                }
    // reductions over the valid region
    for (int c = 0; c < ncomp; ++c) {
        long long sum = 0;
        int mn = 1 << 30, mx = -(1 << 30);
        for (int k = dom.smallEnd(2); k <= dom.bigEnd(2); ++k)
            for (int j = dom.smallEnd(1); j <= dom.bigEnd(1); ++j)
                for (int i = dom.smallEnd(0); i <= dom.bigEnd(0); ++i) {
                    sum += ref(i, j, k, c); mn = std::min(mn, ref(i, j, k, c)); mx = std::max(mx, ref(i, j, k, c));
                }
        if (a.sum(c) != sum || a.min(c) != mn || a.max(c) != mx) ++bad;
    }
    // Copy between fields with different ghost widths, component offsets
    iMultiFab dst(ba, dm, ncomp + 1, ngrow + 1);
    dst.setVal(-7);
    Copy(dst, a, 0, 1, ncomp, ngrow);
    const Box gd = grow(dom, ngrow + 1);
    for (int c = 0; c < ncomp + 1; ++c)
        for (int k = gd.smallEnd(2); k <= gd.bigEnd(2); ++k)
            for (int j = gd.smallEnd(1); j <= gd.bigEnd(1); ++j)
                for (int i = gd.smallEnd(0); i <= gd.bigEnd(0); ++i) {
                    const bool in_g = g.contains(IntVect(i, j, k));
                    const int expect = (c >= 1 && in_g) ? a(i, j, k, c - 1) : -7;
                    if (dst(i, j, k, c) != expect) ++bad;
                }
    return bad;
}

int main() {
    int bad = 0, cases = 0;
    unsigned seed = 1;
    for (int ngrow = 0; ngrow <= 2; ++ngrow)
        for (int per = 0; per < 8; ++per)
            for (int shape = 0; shape < 3; ++shape) {
                const int nx = shape == 0 ? 5 : (shape == 1 ? 1 : 7), ny = shape == 0 ? 4 : 6, nz = shape == 2 ? 2 : 3;
                bad += check(nx, ny, nz, ngrow, 1 + (shape & 1), per & 1, (per >> 1) & 1, (per >> 2) & 1, seed++);
                ++cases;
            }
    std::printf("cases %d mismatches %d\n", cases, bad);
    return bad == 0 ? 0 : 1;
}
