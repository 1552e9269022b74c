// blbm.hpp — C++ host-side mirror of lbm-wgpu's simulation interface over the C ABI (blbm.h).
//
// The reference's host side is Rust (`pub struct LBM`, lbm-wgpu/src/lbm.rs:32-98, and the barrier shapes of
// lbm-wgpu/src/barrier_shapes/); no Rust toolchain exists in this environment, so the compiled-language host
// layer is this header: same type and method names, same argument meaning, same error behaviour (the
// reference panics through expect/unwrap — here blbm::Error is thrown; Line::new returns Result<Line, String> —
// here Line::make returns false and leaves a message).  `const Driver&` survives as an empty struct so call
// sites shaped like lib.rs:46-197 read the same.  Header-only; link with -lblbm.
#pragma once
#include <cstdint>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "blbm.h"

namespace blbm {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("blbm error " + std::to_string(c) + ": " + m), code(c) {}
};

inline void check(int rc)
{
    if (rc != BLBM_OK) throw Error(rc, blbm_last_error());
}

// lbm.rs:10-16, :18-24
enum class SummaryStat { Curl = 0, Ux = 1, Uy = 2, Rho = 3, Speed = 4 };
enum class ColorMap { Inferno = 0, Viridis = 1, Jet = 2 };

// driver.rs:1-7: the wgpu device/queue/surface bundle; the CUDA device and stream live inside the handle
// One device: the whole lattice on it.  Several: the lattice is cut into y-slabs, one per listed device, behind
// the same LBM value (blbm_create_group).
struct Driver {
    std::vector<int> devices{0};
    Driver() {}
    explicit Driver(int device) : devices{device} {}
    explicit Driver(const std::vector<int> &devs) : devices(devs) {}
};

// barrier_shapes/mod.rs:11-19
typedef std::tuple<std::int64_t, std::int64_t, bool> Point;
struct Shape {
    virtual ~Shape() {}
    virtual const std::set<Point> &get_points() const = 0;
    bool is_empty() const { return get_points().empty(); }
};

// barrier_shapes/blob.rs
struct Blob : Shape {
    std::set<Point> points;
    const std::set<Point> &get_points() const override { return points; }
    static Blob new_empty() { return Blob(); }
    void add(const std::vector<Point> &pts, std::uint32_t xdim, std::uint32_t ydim)  // blob.rs:29-36
    {
        for (const Point &p : pts)
            if ((std::uint32_t)std::get<0>(p) < xdim && (std::uint32_t)std::get<1>(p) < ydim) {
                points.erase(Point(std::get<0>(p), std::get<1>(p), !std::get<2>(p)));
                points.insert(p);
            }
    }
    void join(const Shape &s)  // blob.rs:38-43: last writer wins per cell
    {
        for (const Point &p : s.get_points()) {
            points.erase(Point(std::get<0>(p), std::get<1>(p), !std::get<2>(p)));
            points.insert(p);
        }
    }
    void empty() { points.clear(); }
};

// barrier_shapes/line.rs
struct Line : Shape {
    std::set<Point> points;
    const std::set<Point> &get_points() const override { return points; }
    // Line::new / Line::new_erased; false (and *err set) where the reference returns Err
    static bool make(Line *out, std::pair<std::int64_t, std::int64_t> p1, std::pair<std::int64_t, std::int64_t> p2,
                     std::int64_t xdim, std::int64_t ydim, bool erased = false, std::string *err = nullptr)
    {
        std::size_t n = 0;
        int rc = blbm_rasterize_line(p1.first, p1.second, p2.first, p2.second, xdim, ydim, erased, nullptr, 0, &n);
        if (rc != BLBM_OK) {
            if (err)
                *err = "Endpoints (" + std::to_string(p1.first) + "," + std::to_string(p1.second) + ") (" +
                       std::to_string(p2.first) + "," + std::to_string(p2.second) + ") are invalid with dimensions " +
                       std::to_string(xdim) + " and " + std::to_string(ydim);
            return false;
        }
        std::vector<std::int64_t> xy(2 * n);
        blbm_rasterize_line(p1.first, p1.second, p2.first, p2.second, xdim, ydim, erased, xy.data(), n, &n);
        out->points.clear();
        for (std::size_t q = 0; q < n; q++) out->points.insert(Point(xy[2 * q], xy[2 * q + 1], !erased));
        return true;
    }
};

// barrier_shapes/curve.rs
struct Curve : Shape {
    std::set<Point> points;
    bool has_last = false;
    std::pair<std::int64_t, std::int64_t> last_point;
    const std::set<Point> &get_points() const override { return points; }
    void segment(std::pair<std::int64_t, std::int64_t> next, std::int64_t xdim, std::int64_t ydim, bool erased)
    {
        if (has_last) {
            Line l;
            std::string err;
            if (!Line::make(&l, last_point, next, xdim, ydim, erased, &err)) throw Error(BLBM_EINVAL, err);  // .unwrap()
            points.insert(l.points.begin(), l.points.end());
        } else {
            points.insert(Point(next.first, next.second, !erased));
        }
        last_point = next;
        has_last = true;
    }
    void add_segment(std::pair<std::int64_t, std::int64_t> next, std::int64_t xdim, std::int64_t ydim) { segment(next, xdim, ydim, false); }
    void erase_segment(std::pair<std::int64_t, std::int64_t> next, std::int64_t xdim, std::int64_t ydim) { segment(next, xdim, ydim, true); }
    void empty()
    {
        points.clear();
        has_last = false;
    }
};

// lbm.rs:32-98
class LBM {
public:
    ColorMap color_map = ColorMap::Jet;
    std::size_t compute_step = 0;

    // LBM::new(&driver, omega, x, y), lbm.rs:726
    LBM(const Driver &driver, float omega, std::uint32_t x, std::uint32_t y) : x_(x), y_(y)
    {
        check(blbm_create_group(x, y, omega, 0.1f, driver.devices.data(), (int)driver.devices.size(), &h_));
    }
    ~LBM() { blbm_destroy(h_); }
    LBM(const LBM &) = delete;
    LBM &operator=(const LBM &) = delete;

    // lbm.rs:1065: n steps, summary, colour map (render stays with the caller: it needs a surface)
    void iterate(const Driver &, std::size_t compute_steps)
    {
        check(blbm_iterate(h_, (std::uint32_t)compute_steps));
        check(blbm_color_map(h_, (int)color_map));
        compute_step = (std::size_t)blbm_get_compute_num(h_);
    }
    void rerender(const Driver &)  // lbm.rs:1104
    {
        check(blbm_rerender(h_));
        check(blbm_color_map(h_, (int)color_map));
    }
    void collide(const Driver &) { check(blbm_collide(h_)); }                              // lbm.rs:1118
    void stream(const Driver &) { check(blbm_stream(h_)); }                                // lbm.rs:1127
    void set_summary(SummaryStat stat) { check(blbm_set_summary(h_, (int)stat)); }         // lbm.rs:1061
    void update_omega_buffer(const Driver &, float omega) { check(blbm_set_omega(h_, omega)); }  // lbm.rs:1358
    void reset_barrier(const Driver &) { check(blbm_reset_barrier(h_)); }                  // lbm.rs:1362
    void reset_to_equilibrium(const Driver &)                                              // lbm.rs:1076
    {
        check(blbm_reset_to_equilibrium(h_));
        compute_step = 0;
    }
    void custom_speed(const Driver &, float ux)  // lbm.rs:1090
    {
        check(blbm_custom_speed(h_, ux));
        compute_step = 0;
    }
    void single_cell(const Driver &, std::size_t index)  // lbm.rs:1502
    {
        check(blbm_single_cell(h_, (std::uint32_t)index));
        compute_step = 0;
    }
    // lbm.rs:1337 + merge_shapes.rs:12-22
    void draw_shape(const Driver &, const Shape &shape)
    {
        std::vector<std::uint64_t> pairs;
        for (const Point &p : shape.get_points()) {
            pairs.push_back((std::uint64_t)std::get<0>(p) + (std::uint64_t)std::get<1>(p) * x_);
            pairs.push_back(std::get<2>(p) ? 1 : 0);
        }
        if (pairs.empty()) throw Error(BLBM_EINVAL, "empty shape (the reference underflows here, lbm.rs:1343)");
        check(blbm_draw_points64(h_, pairs.data(), pairs.size() / 2));
    }
    void curl_barrier(const Driver &) { check(blbm_curl_barrier(h_)); }        // lbm.rs:1367
    void chaos_barrier(const Driver &) { check(blbm_chaos_barrier(h_)); }      // lbm.rs:1372
    void welcome_barrier(const Driver &) { check(blbm_welcome_barrier(h_)); }  // lbm.rs:1388
    std::size_t get_frame_num() const { return (std::size_t)blbm_get_frame_num(h_); }      // lbm.rs:1166
    std::size_t get_compute_num() const { return (std::size_t)blbm_get_compute_num(h_); }  // lbm.rs:1170

    // ---- new: read back (the reference has no read-back path) ----
    std::vector<float> read_population(int k, int buffer = -1)
    {
        std::vector<float> v((std::size_t)x_ * y_);
        check(blbm_read_population(h_, buffer, k, v.data()));
        return v;
    }
    void read_moments(std::vector<float> *mx, std::vector<float> *my, std::vector<float> *rho)
    {
        const std::size_t n = (std::size_t)x_ * y_;
        mx->resize(n);
        my->resize(n);
        rho->resize(n);
        check(blbm_read_moments(h_, mx->data(), my->data(), rho->data()));
    }
    std::vector<float> read_output()
    {
        std::vector<float> v((std::size_t)x_ * y_);
        check(blbm_read_output(h_, v.data()));
        return v;
    }
    std::vector<float> read_colors()
    {
        std::vector<float> v((std::size_t)x_ * y_ * 3);
        check(blbm_read_colors(h_, v.data()));
        return v;
    }
    std::vector<std::uint32_t> read_barrier()
    {
        std::vector<std::uint32_t> v((std::size_t)x_ * y_);
        check(blbm_read_barrier(h_, v.data()));
        return v;
    }
    blbm_t *handle() { return h_; }
    std::uint32_t x() const { return x_; }
    std::uint32_t y() const { return y_; }

private:
    blbm_t *h_ = nullptr;
    std::uint32_t x_, y_;
};

}  // namespace blbm
