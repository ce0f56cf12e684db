(ns rtclj.native
  "One downcall into librtclj_b200.so (include/rtclj_b200.h) in place of the render loops of
   raytracing/-main and realm.raytracing/-main.  JDK >= 22 (java.lang.foreign).
   UNEXECUTED in the build image (no JVM there); the same ABI is exercised from Python."
  (:import [java.lang.foreign Arena FunctionDescriptor Linker Linker$Option MemoryLayout
            MemorySegment SymbolLookup ValueLayout]))

(def ^:private ^Linker linker (Linker/nativeLinker))
(def ^:private lib (SymbolLookup/libraryLookup "librtclj_b200.so" (Arena/global)))

(defn- downcall [^String sym ^FunctionDescriptor fd]
  (.downcallHandle linker (.orElseThrow (.find lib sym)) fd (make-array Linker$Option 0)))

;; int rtclj_render(const rtclj_scene*, const rtclj_camera*, const rtclj_params*,
;;                  double* out_linear, uint8_t* out_rgb8, rtclj_stats*)
(def ^:private rtclj-render
  (downcall "rtclj_render"
            (FunctionDescriptor/of ValueLayout/JAVA_INT
                                   (into-array MemoryLayout (repeat 6 ValueLayout/ADDRESS)))))
;; const char* rtclj_last_error(void)
(def ^:private rtclj-last-error
  (downcall "rtclj_last_error" (FunctionDescriptor/of ValueLayout/ADDRESS (make-array MemoryLayout 0))))

(def flags-main  (bit-or 1 2 4 8)) ; near-zero guard | Schlick | innermost-first product | sum / spp
(def flags-realm 0)
(def flags-i     (bit-or 16 32))   ; normal shading | int(255.999 c)

(defn- doubles-seg ^MemorySegment [^Arena a xs]
  (.allocateFrom a ValueLayout/JAVA_DOUBLE (double-array xs)))

(defn- last-error ^String []
  (let [^MemorySegment p (.invokeWithArguments rtclj-last-error [])]
    (.getString (.reinterpret p 512) 0)))

(defn render
  "bodies : hittable list made with rtclj.scene, in LIST ORDER (the first body wins a tie).
   cam    : {:pixel-00-loc :pixel-du :pixel-dv :camera-center :defocus-disk-u :defocus-disk-v
            :defocus-angle :image-width :image-height} -- the locals of raytracing.clj:105-139,
            each vector a double[3].
   Returns a vector of double[3] (linear RGB), row-major from the top-left pixel, i.e. `colors`
   of raytracing.clj:170-171, ready for the existing write-color! loop (:172-175)."
  [bodies cam samples-per-px max-depth & {:keys [seed flags device]
                                          :or {seed 1 flags flags-main device 0}}]
  (with-open [a (Arena/ofConfined)]
    (let [n      (count bodies)
          w      (int (:image-width cam))
          h      (int (:image-height cam))
          ;; rtclj_scene (56 bytes): int32 n; int32 pad; 6 pointers
          scene  (doto (.allocate a 56 8)
                   (.set ValueLayout/JAVA_INT 0 (int n))
                   (.set ValueLayout/ADDRESS  8 (doubles-seg a (mapcat :rtclj/center bodies)))
                   (.set ValueLayout/ADDRESS 16 (doubles-seg a (map :rtclj/radius bodies)))
                   (.set ValueLayout/ADDRESS 24 (.allocateFrom a ValueLayout/JAVA_INT
                                                               (int-array (map :rtclj/kind bodies))))
                   (.set ValueLayout/ADDRESS 32 (doubles-seg a (mapcat :rtclj/albedo bodies)))
                   (.set ValueLayout/ADDRESS 40 (doubles-seg a (map :rtclj/fuzz bodies)))
                   (.set ValueLayout/ADDRESS 48 (doubles-seg a (map :rtclj/ior bodies))))
          ;; rtclj_camera (160 bytes): 6 x double[3]; double defocus_angle; int32 width, height
          camera (.allocate a 160 8)
          put3   (fn [^long off ^doubles v]
                   (dotimes [k 3]
                     (.set camera ValueLayout/JAVA_DOUBLE (+ off (* 8 k)) (aget v k))))
          _      (do (put3 0 (:pixel-00-loc cam)) (put3 24 (:pixel-du cam)) (put3 48 (:pixel-dv cam))
                     (put3 72 (:camera-center cam)) (put3 96 (:defocus-disk-u cam))
                     (put3 120 (:defocus-disk-v cam))
                     (.set camera ValueLayout/JAVA_DOUBLE 144 (double (:defocus-angle cam)))
                     (.set camera ValueLayout/JAVA_INT 152 w)
                     (.set camera ValueLayout/JAVA_INT 156 h))
          ;; rtclj_params (48 bytes): spp, depth (int32); seed (uint64); flags (uint32);
          ;; samples_per_unit, shard_index, shard_count, shard_rows, device, pad (int32)
          params (doto (.allocate a 48 8)
                   (.set ValueLayout/JAVA_INT 0 (int samples-per-px))
                   (.set ValueLayout/JAVA_INT 4 (int max-depth))
                   (.set ValueLayout/JAVA_LONG 8 (long seed))
                   (.set ValueLayout/JAVA_INT 16 (int flags))
                   (.set ValueLayout/JAVA_INT 20 (int samples-per-px)) ; the reference's summation order
                   (.set ValueLayout/JAVA_INT 36 (int device)))
          out    (.allocate a (* 8 3 (long w) (long h)) 8)
          rc     (int (.invokeWithArguments rtclj-render
                                            [scene camera params out MemorySegment/NULL MemorySegment/NULL]))]
      (when-not (zero? rc)
        (throw (ex-info (last-error) {:rtclj/code rc})))
      (mapv (fn [^long p]
              (double-array [(.getAtIndex out ValueLayout/JAVA_DOUBLE (* 3 p))
                             (.getAtIndex out ValueLayout/JAVA_DOUBLE (+ 1 (* 3 p)))
                             (.getAtIndex out ValueLayout/JAVA_DOUBLE (+ 2 (* 3 p)))]))
            (range (* (long w) (long h)))))))

(defn render-into-realm!
  "realm.raytracing: copies the linear image into realm[0 .. 3*W*H) where the reference's loop
   (realm/raytracing.clj:325-346) leaves it, so its PPM block (:350-358) can stay."
  [^doubles realm bodies cam samples-per-px max-depth & opts]
  (let [colors (apply render bodies cam samples-per-px max-depth :flags flags-realm opts)]
    (dotimes [p (count colors)]
      (let [^doubles c (nth colors p)]
        (aset realm (* 3 p) (aget c 0))
        (aset realm (+ 1 (* 3 p)) (aget c 1))
        (aset realm (+ 2 (* 3 p)) (aget c 2))))
    realm))

;; int rtclj_encode_ppm_p3_gpu(int32_t device, const uint8_t* rgb8, int32_t width, int32_t height,
;;                             char* out, size_t capacity, size_t* len)
(def ^:private rtclj-encode-ppm-p3-gpu
  (downcall "rtclj_encode_ppm_p3_gpu"
            (FunctionDescriptor/of ValueLayout/JAVA_INT
                                   (into-array MemoryLayout
                                               [ValueLayout/JAVA_INT ValueLayout/ADDRESS ValueLayout/JAVA_INT ValueLayout/JAVA_INT
                                                ValueLayout/ADDRESS ValueLayout/JAVA_LONG ValueLayout/ADDRESS]))))

(defn write-ppm!
  "The write-color! loop (raytracing.clj:172-175) as ONE call: rgb8 is a byte[] of W*H*3 gamma-encoded
   components (what write-color! computes per pixel; rtclj_render fills it when asked for out_rgb8);
   the P3 text is produced by the device kernels and handed to a single .write."
  [^String path ^bytes rgb8 width height & {:keys [device] :or {device 0}}]
  (with-open [a (Arena/ofConfined)]
    (let [w     (int width) h (int height)
          src   (.allocateFrom a ValueLayout/JAVA_BYTE rgb8)
          cap   (+ 64 (* 12 (long w) (long h)))            ; "255 255 255\n" per pixel + header
          out   (.allocate a cap 16)
          len   (.allocate a 8 8)
          rc    (int (.invokeWithArguments rtclj-encode-ppm-p3-gpu [(int device) src w h out cap len]))]
      (when-not (zero? rc)
        (throw (ex-info (last-error) {:rtclj/code rc})))
      (let [n (.get len ValueLayout/JAVA_LONG 0)]
        (with-open [o (java.io.FileOutputStream. path)]
          (.write (.getChannel o) (.asByteBuffer (.asSlice out 0 n))))))))
