(ns rtclj.rng-shim
  "Deterministic draw stream for the JVM reference, so that a `clojure -M:main` / `-M:realm` run can be
   compared draw for draw with librtclj_b200.so and with oracle/rt_oracle.c.

   UNEXECUTED SOURCE: the build image has no JVM.  The hand-out order below is the one of
   oracle/rt_oracle.c (header comment) and DESIGN.md section 2; tests/test_host.py checks a Python
   model of exactly this state machine (tests/rng_shim_model.py) against the oracle's Philox words.

   The reference draws from two unseeded sites: clojure.core/rand (vec3a.clj:72, raytracing.clj:146-147,
   material.clj:42) and realm.rng/rng (realm/rng.clj:6).  Both are replaced by Philox4x32-10 with
     key     = (seed low word, seed high word)
     counter = (pixel index i + j*W, sample k, stage, block)
     stage 0      camera ray: the n-th draw of the sample is word n%4 of block n/4
                  (jitter-x, jitter-y, then the defocus-disk candidates, two words each)
     stage s >= 1 the s-th hit of the path.  Inside random-unit-vec3 the n-th draw is coordinate n%3 of
                  candidate n/3; candidate c is the 64 bits (word 2h, word 2h+1) of block c/2, h = c%2,
                  and a coordinate is a 21-bit field f of it, returned as f * 2^-21.
                  The one draw outside random-unit-vec3 (Schlick, material.clj:42) is word 0 of block 0.
     every draw that is not a 21-bit field is (word >>> 8) * 2^-24.

   Three edits wire it into the reference (they do not change what it computes):
     main   raytracing.clj:143   wrap the body of the k-loop in (rng/with-sample (+ i (* j image-width)) k ...)
            raytracing.clj:45    first form of ray-color: (rng/set-stage! (inc (- max-depth depth)))
                                 (max-depth is a local of -main: pass it in or def it)
            run -main inside     (with-redefs [clojure.core/rand rng/next-uniform
                                               vec3a/random-unit-vec3 (rng/unit-vector-scope vec3a/random-unit-vec3)] ...)
     realm  realm/raytracing.clj:331  (rng/begin-sample! (+ i (* j image-width)) sample) at the top of the sample loop
            realm/raytracing.clj:210  (rng/next-stage!) where Ray.rayColor has found a hit, before .scatter
            once at start-up          (rng/install-realm!)"
  (:import [java.util.random RandomGenerator]))

(set! *unchecked-math* true)

(def ^:const M0 0xD2511F53)
(def ^:const M1 0xCD9E8D57)
(def ^:const W0 0x9E3779B9)
(def ^:const W1 0xBB67AE85)
(def ^:const MASK32 0xFFFFFFFF)

(defn philox4x32-10
  "Counter words c0..c3 and key words k0, k1, all in [0, 2^32) -> long-array of the 4 output words."
  ^longs [^long c0 ^long c1 ^long c2 ^long c3 ^long k0 ^long k1]
  (loop [r 0, c0 c0, c1 c1, c2 c2, c3 c3, k0 k0, k1 k1]
    (if (< r 10)
      (let [p0 (unchecked-multiply (long M0) c0)     ; both factors < 2^32: the 64-bit product is exact
            p1 (unchecked-multiply (long M1) c2)]
        (recur (inc r)
               (bit-and (bit-xor (unsigned-bit-shift-right p1 32) c1 k0) MASK32)
               (bit-and p1 MASK32)
               (bit-and (bit-xor (unsigned-bit-shift-right p0 32) c3 k1) MASK32)
               (bit-and p0 MASK32)
               (bit-and (unchecked-add k0 (long W0)) MASK32)
               (bit-and (unchecked-add k1 (long W1)) MASK32)))
      (long-array [c0 c1 c2 c3]))))

;; Per-thread state (the reference renders on pool threads, raytracing.clj:157-166):
;;   [0] pixel  [1] sample  [2] stage  [3] draws handed out in this stage  [4] 1 inside random-unit-vec3
;;   [5] block number of the cached words, -1 = none
(def ^:private ^ThreadLocal state
  (ThreadLocal/withInitial (reify java.util.function.Supplier (get [_] (long-array [0 0 0 0 0 -1])))))
(def ^:private ^ThreadLocal words
  (ThreadLocal/withInitial (reify java.util.function.Supplier (get [_] (long-array 4)))))
(def seed (atom 1))

(defn- block-words ^longs [^longs st ^long block]
  (let [^longs w (.get words)]
    (when (not= block (aget st 5))
      (let [s   (long @seed)
            out (philox4x32-10 (aget st 0) (aget st 1) (aget st 2) block
                               (bit-and s MASK32) (bit-and (unsigned-bit-shift-right s 32) MASK32))]
        (System/arraycopy out 0 w 0 4)
        (aset st 5 block)))
    w))

(defn begin-sample! [^long pixel ^long sample]
  (let [^longs st (.get state)]
    (aset st 0 pixel) (aset st 1 sample) (aset st 2 0) (aset st 3 0) (aset st 4 0) (aset st 5 -1)))

(defn set-stage! [^long stage]
  (let [^longs st (.get state)]
    (aset st 2 stage) (aset st 3 0) (aset st 5 -1)))

(defn next-stage! []
  (let [^longs st (.get state)] (set-stage! (inc (aget st 2)))))

(defmacro with-sample [pixel sample & body]
  `(do (begin-sample! ~pixel ~sample) ~@body))

(defn next-uniform
  "Drop-in for (rand) / RandomGenerator.nextDouble(): the next draw of the current (pixel, sample, stage)."
  ^double []
  (let [^longs st (.get state)
        n         (aget st 3)]
    (aset st 3 (inc n))
    (cond
      ;; camera ray: plain sequential words
      (zero? (aget st 2))
      (let [w (block-words st (quot n 4))]
        (* (double (unsigned-bit-shift-right (aget w (rem n 4)) 8)) (/ 1.0 16777216.0)))

      ;; inside random-unit-vec3: 21-bit fields of 64-bit candidates
      (== 1 (aget st 4))
      (let [cand  (quot n 3)
            coord (rem n 3)
            w     (block-words st (quot cand 2))
            h     (* 2 (rem cand 2))
            bits  (bit-or (aget w h) (bit-shift-left (aget w (inc h)) 32))   ; 64 bits, high word may set the sign: shifts below are unsigned
            field (bit-and (unsigned-bit-shift-right bits (* 21 coord)) 0x1FFFFF)]
        (* (double field) (/ 1.0 2097152.0)))

      ;; the Schlick draw (material.clj:42): word 0 of block 0, and it is the only draw of its stage
      :else
      (let [w (block-words st 0)]
        (* (double (unsigned-bit-shift-right (aget w 0) 8)) (/ 1.0 16777216.0))))))

(defn unit-vector-scope
  "Wraps vec3a/random-unit-vec3 (vec3a.clj:74-79): its draws are candidate coordinates; the draw
   counter of the stage restarts so that candidate 0 starts at field 0."
  [f]
  (fn [& args]
    (let [^longs st (.get state)]
      (aset st 3 0)
      (aset st 4 1)
      (try (apply f args) (finally (aset st 4 0))))))

(defn install-realm!
  "realm draws only through realm.rng/rng (realm/rng.clj:6-10): jitter via nextDouble(-0.5, 0.5)
   (realm/raytracing.clj:333,335) and candidate coordinates via nextDouble(-1.0, 1.0)
   (realm/vec3.clj:114-116).  The JDK's bounded form is origin + (bound - origin) * nextDouble(), the
   same value as vec3a/rand-double.  realm has no Schlick draw, so every draw of a stage >= 1 is a
   candidate coordinate."
  []
  (alter-var-root (resolve 'realm.rng/rng)
                  (constantly
                   (reify RandomGenerator
                     (nextLong [_] (throw (UnsupportedOperationException. "only nextDouble is used")))
                     (nextDouble [_]
                       (let [^longs st (.get state)]
                         (aset st 4 (if (zero? (aget st 2)) 0 1))
                         (next-uniform)))))))
