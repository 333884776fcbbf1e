// VolumeFraction::value: integer count of fab == phase over the valid cells
// (reference src/props/VolumeFraction.cpp:22-66), executed by the CUDA count
// kernel behind oi_count_phase_i32.  No CPU path: a missing GPU aborts.
#include "VolumeFraction.H"

#include <openimpala_b200.h>

namespace OpenImpala {

VolumeFraction::VolumeFraction(const amrex::iMultiFab& fm, const int phase, int comp)
    : m_mf(fm), m_phase(phase), m_comp(comp) {
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_comp >= 0 && m_comp < m_mf.nComp(),
                                     "VolumeFraction: Component index out of bounds.");
}

void VolumeFraction::value(long long& phase_count, long long& total_count, bool /*local*/) const {
    const std::vector<int> cells = m_mf.validCopy(m_comp);
    int64_t pc = 0, tc = 0;
    const int rc = oi_count_phase_i32(cells.data(), (int64_t)cells.size(), m_phase, &pc, &tc);
    if (rc != OI_OK) amrex::Abort(std::string("VolumeFraction: ") + oi_last_error());
    phase_count = pc;
    total_count = tc;
}

}  // namespace OpenImpala
