#!/bin/bash
# Builds oracle/_ref/petsc_ksp_driver against a PETSc found on THIS box ($PETSC_DIR[/$PETSC_ARCH] or pkg-config).
# Test infrastructure; exits 3 when no PETSc is found (the case in the development image and on the GPU boxes so far).
set -e
here=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$here/_ref"
if pkg-config --exists PETSc 2>/dev/null; then P=PETSc; elif pkg-config --exists petsc 2>/dev/null; then P=petsc; fi
if [ -n "$P" ]; then
  cc=$(pkg-config --variable=ccompiler $P); cc=${cc:-mpicc}
  $cc -O2 $(pkg-config --cflags $P) "$here/petsc_ksp_driver.c" -o "$here/_ref/petsc_ksp_driver" $(pkg-config --libs $P) -lm
elif [ -n "$PETSC_DIR" ] && [ -d "$PETSC_DIR" ]; then
  inc="-I$PETSC_DIR/include"; lib="$PETSC_DIR/lib"
  if [ -n "$PETSC_ARCH" ]; then inc="$inc -I$PETSC_DIR/$PETSC_ARCH/include"; lib="$PETSC_DIR/$PETSC_ARCH/lib"; fi
  mpicc -O2 $inc "$here/petsc_ksp_driver.c" -o "$here/_ref/petsc_ksp_driver" -L"$lib" -Wl,-rpath,"$lib" -lpetsc -lm
else
  echo "no PETSc on this box (PETSC_DIR unset, pkg-config finds none)"; exit 3
fi
echo "$here/_ref/petsc_ksp_driver"
