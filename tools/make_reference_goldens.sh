#!/bin/sh
# Pins the oracle against the REAL reference (needs gfortran; none exists in the build image or on the GPU box,
# so this script is documentation-as-code for a maintainer who has one).
#
# For every shipped example it builds the example's own source copy with an added unformatted dump of
# x, vel, stress, Internal_Vars(1,:) and the ordered pair list after steps 1, 10 and 100, runs it, and compares the
# dumps with the oracle (tools/compare_reference_dump.py). Usage:
#     REF=/path/to/Stress-Particle-SPH sh tools/make_reference_goldens.sh
set -e
REF=${REF:-/root/reference}
command -v gfortran >/dev/null 2>&1 || { echo "gfortran not available: oracle stays unpinned"; exit 0; }
OUT=oracle/_ref
mkdir -p "$OUT"
for ex in soil_failure_bui_et_al_2008 vertical_slope strain_localisation_in_soil_sample; do
  W="$OUT/$ex"; rm -rf "$W"; mkdir -p "$W"
  cp "$REF/example_problems/$ex"/*.f90 "$W/"
  case $ex in
    soil_failure_bui_et_al_2008) cp "$REF/example_problems/$ex/outside_approach/velocity_vector_update"/co_soil.* "$W/";;
    *) cp "$REF/example_problems/$ex"/*.dat "$REF/example_problems/$ex"/*.pts "$W/";;
  esac
  sed 's/^0.00015,2.5,10$/0.00015,2.5,100/' "$REF/example_problems/$ex/input.txt" > "$W/input.txt"
  # dump hook: after "call time_integration" in the main program
  awk '{print} /call time_integration/ {print "   if (itimestep_sph==1 .or. itimestep_sph==10 .or. itimestep_sph==100) call dump_state"}' \
      "$W/1_SPH_2018.f90" > "$W/1_tmp.f90"
  awk '/^CONTAINS/ && !done {print; print "subroutine dump_state"; print "  character(len=32) :: fn"; \
       print "  write(fn,\"(A,I6.6,A)\") \"dump.\", itimestep_sph, \".bin\""; \
       print "  open(77,file=fn,form=\"unformatted\",access=\"stream\")"; \
       print "  write(77) ntotal, ntotal2, x(:,1:ntotal), vel(:,1:ntotal), stress(:,1:ntotal), Internal_Vars(1,1:ntotal)"; \
       print "  close(77)"; print "end subroutine dump_state"; done=1; next} {print}' "$W/1_tmp.f90" > "$W/1_SPH_2018.f90"
  rm "$W/1_tmp.f90"
  ( cd "$W" && gfortran -O3 -o sph 7_variable_types.f90 6_SPH_time_vars_2018.f90 5_SPH_global_vars_2018.f90 \
      4_SPH_material_vars_2018.f90 3_SPH_material_2018.f90 2_SPH_main_2018.f90 1_SPH_2018.f90 && \
    ( ulimit -s unlimited; ./sph > run.log ) )
  echo "$ex: dumps in $W (compare with: python tools/compare_reference_dump.py $W)"
done
