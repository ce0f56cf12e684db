(ns rtclj.scene
  "Data-recording wrappers around the reference's constructors (same names, same arities).
   The reference keeps centre / radius / albedo / fuzz / refraction-index only inside the
   hit-fn / scatter-fn closures (src/hittable.clj:7-9, src/material.clj:13-15,21-23,34-36);
   the native backend needs them as data, so each wrapper calls the ORIGINAL constructor and
   adds the parameters to the map it returns.  UNEXECUTED in the build image (no JVM)."
  (:require [hittable]
            [material]))

(defn sphere [^doubles center ^double radius]
  (assoc (hittable/sphere center radius)
         :rtclj/center (vec center)
         :rtclj/radius radius))

(defn lambertian [^doubles albedo]
  (assoc (material/lambertian albedo)
         :rtclj/kind 0 :rtclj/albedo (vec albedo) :rtclj/fuzz 0.0 :rtclj/ior 1.0))

(defn metal [^doubles albedo ^double fuzz]
  (assoc (material/metal albedo fuzz)
         :rtclj/kind 1 :rtclj/albedo (vec albedo) :rtclj/fuzz fuzz :rtclj/ior 1.0))

(defn dielectric [refraction-index]
  (assoc (material/dielectric refraction-index)
         :rtclj/kind 2 :rtclj/albedo [1.0 1.0 1.0] :rtclj/fuzz 0.0 :rtclj/ior (double refraction-index)))
