// VolumeFraction: integer count of cells == phase over the valid cells
// (reference src/props/VolumeFraction.cpp:22-66), executed by the CUDA count
// kernel behind oi_count_phase_i32.  No CPU path: a missing GPU aborts.
#include "VolumeFraction.H"

#include <openimpala_b200.h>

namespace OpenImpala {

VolumeFraction::VolumeFraction(const amrex::iMultiFab& fm, const int phase, int comp)
    : m_field(&fm), m_phase_id(phase), m_component(comp) {
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_component >= 0 && m_component < m_field->nComp(),
                                     "VolumeFraction: Component index out of bounds.");
}

// local = false: summed over the ranks (reference VolumeFraction.cpp:58-60, ReduceLongSum)
VolumeFraction::Counts VolumeFraction::counts(bool local) const {
    const std::vector<int> cells = m_field->validCopy(m_component);
    int64_t n_phase = 0, n_total = 0;
    if (oi_count_phase_i32(cells.data(), (int64_t)cells.size(), m_phase_id, &n_phase, &n_total) != OI_OK)
        amrex::Abort(std::string("VolumeFraction: ") + oi_last_error());
    Counts c;
    c.phase = n_phase;
    c.total = n_total;
    if (!local && amrex::ParallelDescriptor::NProcs() > 1) {
        long long v[2] = {c.phase, c.total};
        amrex::ParallelDescriptor::ReduceLongSum(v, 2);
        c.phase = v[0]; c.total = v[1];
    }
    return c;
}

void VolumeFraction::value(long long& phase_count, long long& total_count, bool local) const {
    const Counts c = counts(local);
    phase_count = c.phase;
    total_count = c.total;
}

amrex::Real VolumeFraction::value_vf(bool local) const { return counts(local).fraction(); }

}  // namespace OpenImpala
