"""Python model of integration/clojure/src/rtclj/rng_shim.clj plus a sequential-`rand` restatement of the
reference's `main` and `realm` render loops, used to check that the shim's hand-out order reproduces the
counter-based stream of oracle/rt_oracle.c.

Test infrastructure only (pure-Python loops: tiny images).  `ShimStream` mirrors the Clojure state
machine line for line; `render_main` is written the way the reference is -- recursive ray-color
(raytracing.clj:45-58), materials that call `rand` / `random-unit-vec3` (material.clj:13-46,
vec3a.clj:71-86) -- and draws ONLY through `stream.rand()`, i.e. it does not know the counter layout.
"""
import math


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return (c0, c1, c2, c3)


class ShimStream:
    """state = [pixel, sample, stage, draws handed out in this stage, inside-unit-vector flag]"""

    def __init__(self, seed=1):
        self.seed = seed
        self.pixel = self.sample = self.stage = self.n = 0
        self.unit = False
        self.blocks_drawn = 0
        self._cached = (None, None)

    def _words(self, block):
        key = (self.pixel, self.sample, self.stage, block)
        if self._cached[0] != key:
            self._cached = (key, philox4x32_10(*key, self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF))
            self.blocks_drawn += 1
        return self._cached[1]

    def begin_sample(self, pixel, sample):
        self.pixel, self.sample, self.stage, self.n, self.unit = pixel, sample, 0, 0, False

    def set_stage(self, stage):
        self.stage, self.n = stage, 0

    def rand(self):
        n = self.n
        self.n += 1
        if self.stage == 0:                       # camera ray: plain sequential words
            return (self._words(n // 4)[n % 4] >> 8) * (1.0 / 16777216.0)
        if self.unit:                             # 21-bit fields of 64-bit candidates
            cand, coord = divmod(n, 3)
            w = self._words(cand // 2)
            h = 2 * (cand % 2)
            bits = w[h] | (w[h + 1] << 32)
            return ((bits >> (21 * coord)) & 0x1FFFFF) * (1.0 / 2097152.0)
        return (self._words(0)[0] >> 8) * (1.0 / 16777216.0)   # the Schlick draw

    def unit_vector_scope(self, f):
        def wrapped(*a):
            self.n, self.unit = 0, True
            try:
                return f(*a)
            finally:
                self.unit = False
        return wrapped


# ---- vec3a (vec3a.clj:8-101): tuples, left-to-right sums, true division
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def muls(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def mulv(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
def divs(a, s): return (a[0] / s, a[1] / s, a[2] / s)
def neg(a): return (-a[0], -a[1], -a[2])
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def lensq(a): return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]
def unit(a): return divs(a, math.sqrt(lensq(a)))
def reflect(v, n): return sub(v, muls(n, 2.0 * dot(v, n)))


def jmin1(x):
    return x if x != x else (x if x < 1.0 else 1.0)


def refract(uv, n, eta):
    cos_theta = jmin1(dot(neg(uv), n))
    perp = muls(add(uv, muls(n, cos_theta)), eta)
    para = muls(n, -math.sqrt(abs(1.0 - lensq(perp))))
    return add(perp, para)


def render_main(soa, cam, spp, max_depth, seed=1, realm=False):
    """The `main` variant (Schlick, near-zero guard, defocus disk, innermost-first product, / spp) or,
    with realm=True, the `realm` variant (none of the first three, forward product as
    realm/raytracing.clj:205-236 multiplies it, x (1/spp)); sequential draws either way.
    Returns (linear [H][W] of 3-tuples, segments traced, Philox blocks drawn)."""
    center, radius, kind, albedo, fuzz, ior = (x.tolist() for x in soa)
    st = ShimStream(seed)
    rand = st.rand
    segments = [0]

    def rand_double(lo, hi):                       # vec3a.clj:71-72
        return lo + (hi - lo) * rand()

    def random_unit_vec3():                        # vec3a.clj:74-79
        while True:
            x, y, z = rand_double(-1.0, 1.0), rand_double(-1.0, 1.0), rand_double(-1.0, 1.0)
            l2 = x * x + y * y + z * z
            if 1e-160 < l2 <= 1.0:
                return divs((x, y, z), math.sqrt(l2))
    random_unit_vec3 = st.unit_vector_scope(random_unit_vec3)

    def random_in_unit_disk():                     # vec3a.clj:81-86
        while True:
            x, y = rand_double(-1.0, 1.0), rand_double(-1.0, 1.0)
            if x * x + y * y < 1.0:
                return (x, y, 0.0)

    def hit_anything(o, d, t_min, t_max):          # raytracing.clj:33-43 + hittable.clj:7-31
        best, closest = -1, t_max
        for i in range(len(radius)):
            oc = sub(tuple(center[i]), o)
            a, h = lensq(d), dot(d, oc)
            c = lensq(oc) - radius[i] * radius[i]
            disc = h * h - a * c
            if disc < 0.0:
                continue
            sq = math.sqrt(disc)
            root = (h - sq) / a
            if root <= t_min or closest <= root:
                root = (h + sq) / a
                if root <= t_min or closest <= root:
                    continue
            best, closest = i, root
        if best < 0:
            return None
        p = add(o, muls(d, closest))
        outward = divs(sub(p, tuple(center[best])), radius[best])
        front = dot(d, outward) < 0.0
        return best, p, (outward if front else neg(outward)), front

    BLACK = (0.0, 0.0, 0.0)

    def bounce(o, d, depth):
        """One level of ray-color (raytracing.clj:45-58): ("black",) | ("sky", colour) |
        ("scatter", point, direction, attenuation)."""
        st.set_stage(max_depth - depth + 1)        # the shim's hook: draws of this level belong to this hit
        if depth <= 0:
            return ("black",)
        segments[0] += 1
        rec = hit_anything(o, d, 1e-3, math.inf)
        if rec is None:
            y = unit(d)[1]
            a = 0.5 * (y + 1.0)
            return ("sky", add(muls((1.0, 1.0, 1.0), 1.0 - a), muls((0.5, 0.7, 1.0), a)))
        b, p, n, front = rec
        if kind[b] == 0:                           # lambertian, material.clj:13-19 / realm :138-145
            s = add(random_unit_vec3(), n)
            if not realm and abs(s[0]) < 1e-8 and abs(s[1]) < 1e-8 and abs(s[2]) < 1e-8:
                s = n
            return ("scatter", p, s, tuple(albedo[b]))
        if kind[b] == 1:                           # metal, material.clj:21-28 / realm :147-158
            refl = reflect(d, n)
            refl = add(muls(random_unit_vec3(), fuzz[b]), refl)
            if not dot(refl, n) > 0:
                return ("black",)
            return ("scatter", p, refl, tuple(albedo[b]))
        ri = 1.0 / ior[b] if front else ior[b]     # dielectric, material.clj:34-46 / realm :160-177
        u = unit(d)
        cos_theta = jmin1(dot(neg(u), n))
        sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
        if not ri * sin_theta <= 1.0:
            do_reflect = True
        elif realm:                                # realm/raytracing.clj:169-173: no Schlick, no draw
            do_reflect = False
        else:                                      # `or` short-circuits: the draw happens only here
            q = (1.0 - ri) / (1.0 + ri)
            r0 = q * q
            m = 1.0 - cos_theta
            m2 = m * m
            do_reflect = r0 + (1.0 - r0) * (m2 * m2 * m) > rand()
        return ("scatter", p, reflect(u, n) if do_reflect else refract(u, n, ri), (1.0, 1.0, 1.0))

    def ray_color(o, d, depth):                    # main: recursive, innermost product first (:52-53)
        r = bounce(o, d, depth)
        if r[0] == "black":
            return BLACK
        if r[0] == "sky":
            return r[1]
        return mulv(ray_color(r[1], r[2], depth - 1), r[3])

    def ray_color_realm(o, d):                     # realm/raytracing.clj:205-236: iterative, forward product
        throughput, depth = (1.0, 1.0, 1.0), max_depth
        while True:
            r = bounce(o, d, depth)
            if r[0] == "black":
                return BLACK
            if r[0] == "sky":
                return mulv(throughput, r[1])
            _, o, d, att = r
            throughput = mulv(throughput, att)
            depth -= 1

    W, H = cam.width, cam.height
    out = [[None] * W for _ in range(H)]
    for j in range(H):
        for i in range(W):
            acc = (0.0, 0.0, 0.0)
            for k in range(spp):                   # raytracing.clj:141-155
                st.begin_sample(i + j * W, k)
                sample = add(add(cam.pixel00, muls(cam.pixel_du, i + (rand() - 0.5))),
                             muls(cam.pixel_dv, j + (rand() - 0.5)))
                if cam.defocus_angle <= 0:
                    origin = cam.center
                else:                              # raytracing.clj:89-93
                    pd = random_in_unit_disk()
                    origin = add(add(cam.center, muls(cam.defocus_u, pd[0])), muls(cam.defocus_v, pd[1]))
                if realm:
                    acc = add(acc, ray_color_realm(origin, sub(sample, origin)))
                else:
                    acc = add(acc, ray_color(origin, sub(sample, origin), max_depth))
            out[j][i] = muls(acc, 1.0 / spp) if realm else divs(acc, spp)
    return out, segments[0], st.blocks_drawn
