// test_lbm.cpp — the parity test in the reference's own (compiled) idiom: drive blbm::LBM (include/blbm.hpp,
// the C++ mirror of lbm-wgpu's `pub struct LBM`) and the CPU oracle with the same call sequence and compare
// every buffer bit for bit.  The oracle is linked as the checker only.
//   test_lbm --no-gpu : checks that constructing an LBM without a GPU fails loudly (BLBM_ENOGPU)
//   test_lbm          : needs a GPU
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "blbm.hpp"

extern "C" {
struct lbm_oracle;
lbm_oracle *lbm_oracle_create(uint32_t w, uint32_t h, float omega, float inflow_ux);
void lbm_oracle_destroy(lbm_oracle *o);
void lbm_oracle_iterate(lbm_oracle *o, uint32_t n);
void lbm_oracle_draw_points(lbm_oracle *o, const uint32_t *pairs, size_t npairs);
void lbm_oracle_set_omega(lbm_oracle *o, float omega);
void lbm_oracle_custom_speed(lbm_oracle *o, float ux);
void lbm_oracle_reset_barrier(lbm_oracle *o);
void lbm_oracle_set_summary(lbm_oracle *o, int stat);
float *lbm_oracle_population(lbm_oracle *o, int buffer, int k);
uint32_t *lbm_oracle_barrier(lbm_oracle *o);
float *lbm_oracle_output(lbm_oracle *o);
void lbm_oracle_color_map(const lbm_oracle *o, int map, float *rgb);
uint64_t lbm_oracle_compute_num(const lbm_oracle *o);
}

static int failures = 0;

static void expect_bits(const std::vector<float> &got, const float *want, const char *what)
{
    size_t bad = 0;
    for (size_t i = 0; i < got.size(); i++) {
        uint32_t a, b;
        memcpy(&a, &got[i], 4);
        memcpy(&b, &want[i], 4);
        if (a != b && !(std::isnan(got[i]) && std::isnan(want[i]))) bad++;
    }
    if (bad) {
        printf("FAIL %s: %zu of %zu values differ\n", what, bad, got.size());
        failures++;
    }
}

static void run(const blbm::Driver &driver, const char *label);

int main(int argc, char **argv)
{
    using namespace blbm;
    if (argc > 1 && std::string(argv[1]) == "--no-gpu") {
        try {
            const Driver driver;
            LBM lbm(driver, 1.25f, 64, 32);
        } catch (const Error &e) {
            printf("constructing without a GPU failed as it must: %s\n", e.what());
            return e.code == BLBM_ENOGPU ? 0 : 2;
        }
        printf("a GPU is present\n");
        return 0;
    }
    run(Driver(), "one device");
    // the same LBM value over three y-slabs (blbm_create_group): distinct GPUs when the box has them, else three
    // slabs on device 0 — either way the halo rows travel by direct stores from the step kernel
    const int ndev = blbm_device_count();
    std::vector<int> devs;
    for (int q = 0; q < 3; q++) devs.push_back(ndev >= 3 ? q : 0);
    run(Driver(devs), "three slabs");
    printf(failures ? "FAILED (%d)\n" : "ok: blbm::LBM bit-identical to the oracle (%d failures)\n", failures);
    return failures ? 1 : 0;
}

static void run(const blbm::Driver &driver, const char *label)
{
    using namespace blbm;
    const uint32_t x = 320, y = 96;
    const float omega = 1.0f / (3.0f * 0.02f + 0.5f);  // lib.rs:183
    LBM lbm(driver, omega, x, y);
    lbm_oracle *ora = lbm_oracle_create(x, y, omega, 0.1f);

    auto both_draw = [&](const Shape &s) {
        lbm.draw_shape(driver, s);
        std::vector<uint32_t> pairs;  // get_points_vector, merge_shapes.rs:12-22
        for (const Point &p : s.get_points()) {
            pairs.push_back((uint32_t)std::get<0>(p) + (uint32_t)std::get<1>(p) * x);
            pairs.push_back(std::get<2>(p) ? 1u : 0u);
        }
        lbm_oracle_draw_points(ora, pairs.data(), pairs.size() / 2);
    };
    auto compare = [&](const char *what) {
        const std::string tag = std::string(label) + ": " + what;
        for (int b = 0; b < 2; b++)
            for (int k = 0; k < 9; k++)
                expect_bits(lbm.read_population(k, b), lbm_oracle_population(ora, k == 4 ? 0 : b, k),
                            (tag + " population " + std::to_string(b) + "/" + std::to_string(k)).c_str());
        expect_bits(lbm.read_output(), lbm_oracle_output(ora), (tag + " output").c_str());
        std::vector<float> rgb((size_t)x * y * 3);
        lbm_oracle_color_map(ora, (int)lbm.color_map, rgb.data());
        expect_bits(lbm.read_colors(), rgb.data(), (tag + " colours").c_str());
        const std::vector<uint32_t> bar = lbm.read_barrier();
        if (memcmp(bar.data(), lbm_oracle_barrier(ora), bar.size() * 4) != 0) {
            printf("FAIL %s barrier\n", tag.c_str());
            failures++;
        }
        if (lbm.get_compute_num() != lbm_oracle_compute_num(ora)) {
            printf("FAIL %s compute_num\n", tag.c_str());
            failures++;
        }
    };

    // the call pattern of lib.rs:108-199: paint, iterate(15), change viscosity, erase, switch output
    Line l1, l2;
    std::string err;
    if (!Line::make(&l1, {60, 20}, {60, 70}, x, y) || !Line::make(&l2, {120, 30}, {200, 60}, x, y)) {
        failures++;
        return;
    }
    if (Line::make(&l1, {60, 20}, {(int64_t)x, 70}, x, y, false, &err)) failures++;  // Err, like Line::new
    Line::make(&l1, {60, 20}, {60, 70}, x, y);
    both_draw(l1);
    for (int frame = 0; frame < 20; frame++) {
        lbm.iterate(driver, 15);
        lbm_oracle_iterate(ora, 15);
    }
    compare("after 300 steps");
    Blob blob = Blob::new_empty();
    blob.join(l2);
    Curve eraser;
    eraser.erase_segment({60, 40}, x, y);
    eraser.erase_segment({70, 50}, x, y);
    blob.join(eraser);
    both_draw(blob);
    lbm.update_omega_buffer(driver, 1.6f);
    lbm_oracle_set_omega(ora, 1.6f);
    lbm.set_summary(SummaryStat::Speed);
    lbm_oracle_set_summary(ora, 4);
    lbm.color_map = ColorMap::Viridis;
    for (int frame = 0; frame < 10; frame++) {
        lbm.iterate(driver, 15);
        lbm_oracle_iterate(ora, 15);
    }
    compare("after paint/erase/omega");
    lbm.custom_speed(driver, 0.05f);
    lbm_oracle_custom_speed(ora, 0.05f);
    lbm.reset_barrier(driver);
    lbm_oracle_reset_barrier(ora);
    lbm.iterate(driver, 40);
    lbm_oracle_iterate(ora, 40);
    compare("after custom_speed/reset_barrier");
    lbm_oracle_destroy(ora);
}
