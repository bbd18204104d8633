#!/usr/bin/env bash
# oracle/build_ref.sh -- build oracle/_ref/libalva_ref.so from the reference's own sources.
# TEST INFRASTRUCTURE: the .so is git-ignored, travels to the GPU box, and is only ever loaded by
# tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arms.
#
# Nothing from /root/reference is copied into the repo: the vendored OpenCV 4.5.5 / Ceres 2.0 are
# configured out-of-tree (SURVEY.md Appendix A recipe) into $ALVA_REF_PREFIX (default /tmp/probe; the
# survey's build is reused when present) and AlvaAR's ceres_parametrization.cpp is compiled in place.
# NOTE (DESIGN.md "Oracle"): OpenCV and Ceres cannot be compiled without their generated headers, so
# this step runs their cmake configure offline; it is optional -- every parity test also runs
# against the committed golden vectors with the plain-C restatement when the .so is absent.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${ALVA_REFERENCE:-/root/reference}
P=${ALVA_REF_PREFIX:-/tmp/probe}
OUT="$HERE/_ref"
# ALVA_REF_VARIANT=fma builds a SECOND copy of the reference, libalva_ref_fma.so, with FMA contraction enabled for AlvaAR's own
# sources and OpenGV (-O2 -mfma -ffp-contract=fast: what gcc does by default on aarch64, and what -march=native does on x86;
# -U__AVX__ -U__FMA__ keeps Eigen on the SSE2 packet types the prebuilt Ceres / OpenCV objects use -- only the contraction changes).
# Same sources, another legitimate build: tools/make_golden_system.py stores how far the two builds' trajectories are apart
# (ref_alt_*), the yardstick for the free-running pose tolerance.  Never shipped, never timed.
VARIANT=${ALVA_REF_VARIANT:-}
XFLAGS=""; LIBNAME=libalva_ref.so; GVDIR=opengv
if [ "$VARIANT" = fma ]; then XFLAGS="-mfma -ffp-contract=fast -U__AVX__ -U__FMA__"; LIBNAME=libalva_ref_fma.so; GVDIR=opengv_fma; fi
[ -d "$REF/src/slam/src" ] || { echo "reference tree not found at $REF" >&2; exit 3; }
mkdir -p "$OUT" "$P"
J=${JOBS:-$(nproc)}
if [ ! -f "$P/ocv_install/lib/libopencv_core.a" ]; then
  mkdir -p "$P/ocv" && cd "$P/ocv"
  cmake -G Ninja "$REF/src/libs/opencv" -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DCMAKE_INSTALL_PREFIX="$P/ocv_install" -DBUILD_LIST=core,imgproc,features2d,flann,video,calib3d \
    -DBUILD_SHARED_LIBS=OFF -DBUILD_TESTS=OFF -DBUILD_PERF_TESTS=OFF -DBUILD_EXAMPLES=OFF -DBUILD_opencv_apps=OFF \
    -DBUILD_JAVA=OFF -DBUILD_opencv_python3=OFF -DWITH_IPP=OFF -DWITH_ITT=OFF -DWITH_OPENCL=OFF -DWITH_CUDA=OFF \
    -DWITH_PROTOBUF=OFF -DWITH_QUIRC=OFF -DWITH_ADE=OFF -DWITH_JPEG=OFF -DWITH_PNG=OFF -DWITH_TIFF=OFF -DWITH_WEBP=OFF \
    -DWITH_OPENJPEG=OFF -DWITH_JASPER=OFF -DWITH_OPENEXR=OFF -DWITH_FFMPEG=OFF -DWITH_GSTREAMER=OFF -DWITH_V4L=OFF \
    -DWITH_GTK=OFF -DWITH_EIGEN=OFF -DWITH_LAPACK=OFF -DWITH_1394=OFF -DWITH_VTK=OFF -DBUILD_ZLIB=ON \
    -DCMAKE_POSITION_INDEPENDENT_CODE=ON > cfg.log 2>&1
  ninja -j"$J" install > build.log 2>&1
fi
if [ ! -f "$P/ceres_install/lib/libceres.a" ]; then
  mkdir -p "$P/eigen_stub" "$P/ceres"
  cat > "$P/eigen_stub/Eigen3Config.cmake" <<EOS
if(NOT TARGET Eigen3::Eigen)
  add_library(Eigen3::Eigen INTERFACE IMPORTED)
  set_target_properties(Eigen3::Eigen PROPERTIES INTERFACE_INCLUDE_DIRECTORIES "$REF/src/libs/eigen")
endif()
set(EIGEN3_INCLUDE_DIR "$REF/src/libs/eigen")
set(EIGEN3_INCLUDE_DIRS "$REF/src/libs/eigen")
set(EIGEN3_VERSION_STRING "3.4.0")
set(Eigen3_VERSION "3.4.0")
set(EIGEN3_FOUND TRUE)
EOS
  printf 'set(PACKAGE_VERSION "3.4.0")\nset(PACKAGE_VERSION_COMPATIBLE TRUE)\n' > "$P/eigen_stub/Eigen3ConfigVersion.cmake"
  cd "$P/ceres"
  cmake -G Ninja "$REF/src/libs/ceres-solver" -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
    -DCMAKE_INSTALL_PREFIX="$P/ceres_install" -DEigen3_DIR="$P/eigen_stub" -DBUILD_SHARED_LIBS=OFF -DBUILD_EXAMPLES=OFF \
    -DBUILD_TESTING=OFF -DBUILD_BENCHMARKS=OFF -DEIGENSPARSE=ON -DSUITESPARSE=OFF -DCXSPARSE=OFF -DLAPACK=OFF \
    -DMINIGLOG=ON -DGFLAGS=OFF -DCERES_THREADING_MODEL=NO_THREADS -DCMAKE_POSITION_INDEPENDENT_CODE=ON > cfg.log 2>&1
  ninja -j"$J" install > build.log 2>&1
fi
# OpenGV 1.0 (P3P-Kneip / LMedS for the per-frame pose): its own CMake flags (-march=native -O3) produce a library that
# crashes under gcc 13, so the sources are compiled directly, in place, with plain -O2.
if [ ! -f "$P/$GVDIR/libopengv.a" ]; then
  mkdir -p "$P/$GVDIR/obj"
  GV="$REF/src/libs/opengv"
  find "$GV/src" -name '*.cpp' | grep -v -i -e python -e matlab | while read -r f; do
    o="$P/$GVDIR/obj/$(echo "${f#$GV/src/}" | tr '/' '_' | sed 's/\.cpp$/.o/')"
    echo "g++ -std=c++17 -O2 $XFLAGS -w -fPIC -fno-strict-aliasing -I$GV/include -I$REF/src/libs/eigen -I$REF/src/libs/eigen/unsupported -c $f -o $o"
  done | xargs -P "$J" -I{} sh -c '{}'
  ar rcs "$P/$GVDIR/libopengv.a" "$P"/$GVDIR/obj/*.o
fi
INC="-I$REF/src/slam/src -I$REF/src/libs/opencv/modules/highgui/include -I$REF/src/libs/opencv/modules/imgcodecs/include -I$REF/src/libs/opencv/modules/videoio/include -I$REF/src/libs/opengv/include -I$P/ocv_install/include/opencv4 -I$REF/src/libs/eigen -I$REF/src/libs/Sophus \
 -I$P/ceres_install/include -I$P/ceres_install/include/ceres/internal/miniglog"
cd "$OUT"
# every AlvaAR source except the Emscripten binding, compiled where it lies, unmodified (system.cpp needs C++20 for its
# unqualified duration_cast) -- with ONE documented exception (SURVEY 8c / Appendix B "solver time caps"):
# the three wall-clock caps of the Ceres solves
#     optimizer.cpp:258            options.max_solver_time_in_seconds = 0.01;   (local BA, first solve)
#     optimizer.cpp:322            options.max_solver_time_in_seconds = 0.001;  (local BA, second solve)
#     multi_view_geometry.cpp:185  options.max_solver_time_in_seconds = 0.005;  (PnP)
# make the reference's result depend on the load of the host it runs on.  For those two files a build-time copy under
# _ref/patched/ (git-ignored, generated by the sed below, never committed) routes the literal through alva_ref_time_cap(),
# defined in ref_system.cpp: it returns the literal unchanged by default -- timing arms and every other caller see the
# reference's own behaviour bit for bit -- and 1e9 after ref_config_time_caps(1), which only the golden generators call.
# The whole diff is those three right-hand sides plus one declaration line per file.
mkdir -p "$OUT/patched"
for f in optimizer multi_view_geometry; do
  { echo 'double alva_ref_time_cap(double seconds);';
    sed -E 's/(max_solver_time_in_seconds[[:space:]]*=[[:space:]]*)([0-9.]+)[[:space:]]*;/\1alva_ref_time_cap(\2);/' "$REF/src/slam/src/$f.cpp"; } > "$OUT/patched/$f.cpp"
  n=$(grep -c 'alva_ref_time_cap(' "$OUT/patched/$f.cpp")
  [ "$f" = optimizer ] && want=3 || want=2
  [ "$n" -eq "$want" ] || { echo "time-cap patch of $f.cpp: expected $want matches, got $n" >&2; exit 4; }
done
OBJS=""
for f in camera_calibration ceres_parametrization feature_extractor feature_tracker frame map_manager map_point mapper \
         multi_view_geometry optimizer state system utils visual_frontend; do
  src="$REF/src/slam/src/$f.cpp"
  [ -f "$OUT/patched/$f.cpp" ] && src="$OUT/patched/$f.cpp"
  echo "g++ -std=c++20 -O2 $XFLAGS -w -fPIC -c $src -o $OUT/$f.o $INC"
  OBJS="$OBJS $f.o"
done | xargs -P "$J" -I{} sh -c '{}'
for f in camera_calibration ceres_parametrization feature_extractor feature_tracker frame map_manager map_point mapper \
         multi_view_geometry optimizer state system utils visual_frontend; do OBJS="$OBJS $f.o"; done
g++ -std=c++17 -O2 -w -fPIC -c "$HERE/ref_harness.cpp" -o ref_harness.o $INC
g++ -std=c++20 -O2 -w -fPIC -c "$HERE/ref_system.cpp" -o ref_system.o $INC
g++ -shared -o $LIBNAME ref_harness.o ref_system.o $OBJS \
  -Wl,--start-group "$P"/ocv_install/lib/libopencv_video.a "$P"/ocv_install/lib/libopencv_calib3d.a \
  "$P"/ocv_install/lib/libopencv_features2d.a "$P"/ocv_install/lib/libopencv_flann.a \
  "$P"/ocv_install/lib/libopencv_imgproc.a "$P"/ocv_install/lib/libopencv_core.a \
  "$P"/ocv_install/lib/opencv4/3rdparty/libzlib.a "$P"/ceres_install/lib/libceres.a "$P"/$GVDIR/libopengv.a -Wl,--end-group \
  -lpthread -ldl -static-libstdc++ -static-libgcc -Wl,--exclude-libs,ALL
rm -f ref_harness.o ref_system.o $OBJS
rm -rf "$OUT/patched"
echo "built $OUT/$LIBNAME"
