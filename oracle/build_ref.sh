#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference sources where they lie
# (/root/reference, read-only) into oracle/_ref/ (git-ignored).  No reference source is
# copied into this repository: gcc reads the .c files in place and only object files,
# a generated version header and the asm shim object end up under oracle/_ref/.
#
# The reference's own build (cmake + NASM) is not usable here (no yasm/nasm); the six
# .asm files export 59 symbols, none on the hot path; oracle/ref_asm_shim.c forwards
# each to its plain-C twin that already lives in the reference objects.
set -e
REF=${SVT_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/Source/Lib/Codec" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p $OUT/obj $OUT/inc
sed 's/@SVT_VP9_VERSION_MAJOR@/0/;s/@SVT_VP9_VERSION_MINOR@/3/;s/@SVT_VP9_VERSION_PATCHLEVEL@/1/' \
    $REF/Source/API/EbApiVersion.h.in > $OUT/inc/EbApiVersion.h
L=$REF/Source/Lib
INC="-I$OUT/inc -I$REF/Source/API -I$L/VPX -I$L/Codec -I$L/C_DEFAULT -I$L/ASM_SSE2 -I$L/ASM_SSSE3 -I$L/ASM_SSE4_1 -I$L/ASM_AVX2"
CF="-std=gnu99 -O2 -fPIC -w -g"
echo "$INC" > $OUT/inc_flags.txt
cc_one() { # src flags prefix
  local o=$OUT/obj/$3_$(basename $1 .c).o
  if [ ! -f $o ] || [ $1 -nt $o ]; then gcc $CF $2 $INC -c $1 -o $o; fi
}
export -f cc_one; export OUT CF INC
J=$(nproc)
ls $L/VPX/*.c       | xargs -P $J -I{} bash -c 'cc_one {} "" VPX'
ls $L/Codec/*.c     | xargs -P $J -I{} bash -c 'cc_one {} "" Codec'
ls $L/C_DEFAULT/*.c | xargs -P $J -I{} bash -c 'cc_one {} "" C'
ls $L/ASM_SSE2/*.c  | xargs -P $J -I{} bash -c 'cc_one {} "-msse2" SSE2'
ls $L/ASM_SSSE3/*.c | xargs -P $J -I{} bash -c 'cc_one {} "-mssse3" SSSE3'
ls $L/ASM_SSE4_1/*.c| xargs -P $J -I{} bash -c 'cc_one {} "-msse4.1" SSE41'
ls $L/ASM_AVX2/*.c  | xargs -P $J -I{} bash -c 'cc_one {} "-mavx2" AVX2'
gcc $CF $INC -c $HERE/ref_asm_shim.c -o $OUT/obj/shim.o
# every symbol stays visible (no version script): the tests call internal _c kernels by name
gcc -shared -o $OUT/libSvtVp9Enc_ref.so $OUT/obj/*.o -lpthread -lm -Wl,--no-undefined
gcc -std=gnu99 -O2 -w -I$OUT/inc -I$REF/Source/API $REF/Source/App/*.c -o $OUT/SvtVp9EncApp_ref \
    -L$OUT -lSvtVp9Enc_ref -lpthread -lm -lrt -Wl,-rpath,'$ORIGIN'
echo "build_ref: ok -> $OUT"

# ---- seam-check flavour: same reference objects + a --wrap interposer that runs the oracle port
# next to the reference's motion_estimate_sb (pins oracle/me_port.c against the reference)
gcc $CF $INC -I$HERE/../include -I$HERE/../svt-vp9_b200/shim -c $HERE/ref_seam_check.c -o $OUT/seam_check.o
gcc -std=gnu99 -O2 -fPIC -I$HERE/../include -c $HERE/me_port.c -o $OUT/me_port.o
gcc -std=gnu99 -O2 -fPIC -I$HERE/../include -c $HERE/lf_port.c -o $OUT/lf_port.o
gcc -std=gnu99 -O2 -fPIC -I$HERE/../include -c $HERE/pa_port.c -o $OUT/pa_port.o
gcc -std=gnu99 -O2 -fPIC -I$HERE/../include -c $HERE/pa_noise_port.c -o $OUT/pa_noise_port.o
gcc -shared -o $OUT/libSvtVp9Enc_seam.so $OUT/obj/*.o $OUT/seam_check.o $OUT/me_port.o $OUT/lf_port.o $OUT/pa_port.o $OUT/pa_noise_port.o -lpthread -lm \
    -Wl,--no-undefined -Wl,--wrap=motion_estimate_sb -Wl,--wrap=eb_vp9_loop_filter_frame
gcc -std=gnu99 -O2 -w -I$OUT/inc -I$REF/Source/API $REF/Source/App/*.c -o $OUT/SvtVp9EncApp_seam \
    -L$OUT -lSvtVp9Enc_seam -lpthread -lm -lrt -Wl,-rpath,'$ORIGIN'
echo "build_ref: seam-check flavour ok"

# ---- path driver: the reference's own hot-path functions callable from plain data (bench reference arm + tests)
gcc $CF -fcommon $INC -I$HERE/../include -I$HERE/../svt-vp9_b200/shim -c $HERE/ref_path_driver.c -o $OUT/ref_path_driver.o
gcc -shared -o $OUT/libSvtVp9Enc_refdrv.so $OUT/obj/*.o $OUT/ref_path_driver.o -lpthread -lm -Wl,--no-undefined
echo "build_ref: path driver ok"

# ---- drop-in flavour: the same reference objects with the seam functions bound to libsvtvp9_b200.so (S1 / S2 always, S3 with SVT_VP9_GPU_ENCODE_PASS=1)
# (svt-vp9_b200/shim/svtb200_wrap.c); used by the encoder-level parity test on the GPU box.
# Shape of the reference's own library (Source/Lib/CMakeLists.txt / CMakeLists.txt:61): SONAME libSvtVp9Enc.so.1, symbol map
# Source/API/EbApi.version (SVT_VP9_0 { global: Eb*; eb*; local: *; }) -- read where it lies at link time, not copied.
# eb_vp9_deinit_encoder is provided by the shim (it drops the GPU state first); the reference's definition stays in the link
# under another name: objcopy renames the symbol in a COPY of the compiled object, the source is untouched.
if [ -f $HERE/../svt-vp9_b200/libsvtvp9_b200.so ]; then
  D=$OUT/dropin; mkdir -p $D/obj
  for o in $OUT/obj/*.o; do
    case $(basename $o) in
      Codec_EbEncHandle.o) objcopy --redefine-sym eb_vp9_deinit_encoder=svtb200_real_eb_vp9_deinit_encoder $o $D/obj/$(basename $o) ;;
      *) ln -f $o $D/obj/$(basename $o) 2>/dev/null || cp $o $D/obj/$(basename $o) ;;
    esac
  done
  gcc $CF $INC -I$HERE/../include -I$HERE/../svt-vp9_b200/shim -c $HERE/../svt-vp9_b200/shim/svtb200_wrap.c -o $OUT/svtb200_wrap.o
  gcc $CF $INC -I$HERE/../include -I$HERE/../svt-vp9_b200/shim -c $HERE/../svt-vp9_b200/shim/svtb200_wrap_tx.c -o $OUT/svtb200_wrap_tx.o
  gcc -shared -o $D/libSvtVp9Enc.so.1 $D/obj/*.o $OUT/svtb200_wrap.o $OUT/svtb200_wrap_tx.o -L$HERE/../svt-vp9_b200 -lsvtvp9_b200 -lpthread -lm \
      -Wl,-soname,libSvtVp9Enc.so.1 -Wl,--version-script=$REF/Source/API/EbApi.version \
      -Wl,--wrap=motion_estimate_sb -Wl,--wrap=eb_vp9_loop_filter_frame -Wl,--wrap=eb_vp9_post_full_object -Wl,--wrap=eb_vp9_picture_analysis_context_ctor \
      -Wl,--wrap=eb_vp9_idct32x32_add -Wl,--wrap=eb_vp9_iht16x16_add -Wl,--wrap=eb_vp9_iht8x8_add -Wl,--wrap=eb_vp9_idct4x4_add \
      -Wl,-rpath,'$ORIGIN/../../../svt-vp9_b200'
  gcc -std=gnu99 -O2 -w -I$OUT/inc -I$REF/Source/API $REF/Source/App/*.c -o $OUT/SvtVp9EncApp_b200 \
      -L$D -l:libSvtVp9Enc.so.1 -L$HERE/../svt-vp9_b200 -lsvtvp9_b200 -lpthread -lm -lrt -Wl,-rpath,'$ORIGIN/dropin' -Wl,-rpath,'$ORIGIN/../../svt-vp9_b200'
  rm -rf $D/obj $OUT/libSvtVp9Enc_b200.so
  echo "build_ref: drop-in flavour ok"
fi
