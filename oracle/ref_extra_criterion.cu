// ref_extra_criterion.cu -- TEST INFRASTRUCTURE.  A flow criterion of OUR OWN that is not an index box, handed to the
// reference's unmodified generate_flow kernel through the reference's own mechanism (a __device__ function-pointer
// symbol, FlowSurface.h:15-28): the parity anchor for lbm_set_source_mask.  The same predicate is restated in
// oracle/lbm_oracle.c (LBO_SRC_LATTICE_PATTERN) and in the tests.
namespace LbmExtraCriterion
{
    __device__ bool latticePattern(int x, int y, int z)
    {
        return (x + 2 * y + 3 * z) % 5 == 0 && x >= 1;
    }

    __device__ bool (*platticePattern)(int, int, int) = &LbmExtraCriterion::latticePattern;
}
