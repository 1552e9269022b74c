// curve_chain.cpp — drives blbm::Curve (include/blbm.hpp, the C++ mirror of barrier_shapes/curve.rs) with a stroke
// given on the command line and prints the resulting points, for comparison with the same calls executed in the
// reference's shipped binary (tests/test_wasm_pin.py).  Host code only: runs without a GPU.
//   curve_chain XDIM YDIM  ERASE X Y  [ERASE X Y ...]   ->   lines "x y flag", sorted
#include <cstdio>
#include <cstdlib>

#include "blbm.hpp"

int main(int argc, char **argv)
{
    if (argc < 6 || (argc - 3) % 3 != 0) return 2;
    const long xdim = atol(argv[1]), ydim = atol(argv[2]);
    blbm::Curve c;
    try {
        for (int a = 3; a + 2 < argc; a += 3) {
            const std::pair<std::int64_t, std::int64_t> p(atol(argv[a + 1]), atol(argv[a + 2]));
            if (atoi(argv[a])) c.erase_segment(p, xdim, ydim);
            else c.add_segment(p, xdim, ydim);
        }
    } catch (const blbm::Error &e) {
        printf("error %s\n", e.what());
        return 1;
    }
    for (const blbm::Point &p : c.get_points())
        printf("%lld %lld %d\n", (long long)std::get<0>(p), (long long)std::get<1>(p), (int)std::get<2>(p));
    return 0;
}
