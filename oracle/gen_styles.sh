#!/bin/sh
# oracle/gen_styles.sh <dir> -- TEST INFRASTRUCTURE.
# Writes the style_<kind>.h include lists LAMMPS' factories expect, one "#include" per header in
# <dir> that carries the matching *_CLASS guard macro (the guard names are the ones tested in
# the reference's headers, e.g. pair_ssa_tsdpd_bvf_transport_velocity.h:14 "#ifdef PAIR_CLASS").
set -e
cd "$1"
emit () {   # macro  filename-prefix  kind
  out="style_$3.h"
  : > "$out"
  for f in $(grep -sl "$1" "$2"*.h | grep -v '^style_' | sort); do
    echo "#include \"$f\"" >> "$out"
  done
}
emit ANGLE_CLASS     angle_      angle
emit ATOM_CLASS      atom_vec_   atom
emit BODY_CLASS      body_       body
emit BOND_CLASS      bond_       bond
emit COMMAND_CLASS   ""          command
emit COMPUTE_CLASS   compute_    compute
emit DIHEDRAL_CLASS  dihedral_   dihedral
emit DUMP_CLASS      dump_       dump
emit FIX_CLASS       fix_        fix
emit IMPROPER_CLASS  improper_   improper
emit INTEGRATE_CLASS ""          integrate
emit KSPACE_CLASS    ""          kspace
emit MINIMIZE_CLASS  min_        minimize
emit NBIN_CLASS      nbin_       nbin
emit NPAIR_CLASS     npair_      npair
emit NSTENCIL_CLASS  nstencil_   nstencil
emit NTOPO_CLASS     ntopo_      ntopo
emit PAIR_CLASS      pair_       pair
emit READER_CLASS    reader_     reader
emit REGION_CLASS    region_     region
