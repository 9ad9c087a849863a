! Runs the reference's GJKEPA over a pair file and writes its six outputs (ref_patch recipe; UNTESTED here).
! Input  (stream, little endian):  int64 n_hulls, int64 n_vert, int64 n_pairs, int64 hull_off(n_hulls+1),
!                                  real64 verts(3*n_vert) [hull blocks x(n)|y(n)|z(n)], int32 pair_a(n_pairs), int32 pair_b(n_pairs)
!         = what tools or tests write with numpy:  header.tofile; hull_off.tofile; verts.tofile; pair_a.tofile; pair_b.tofile
! Output (stream): int8 collision(n_pairs), int32 colliType(n_pairs), real64 nearest(6,n_pairs), normal(3,n_pairs),
!                  point(3,n_pairs), depth(n_pairs)      -- compare with oracle/orc.py batch() on the same arrays.
PROGRAM ref_driver
    USE GCLIB_GJKEPA
    IMPLICIT NONE
    CHARACTER(LEN=1024) :: fin, fout
    INTEGER*8 :: nh, nv, np, i
    INTEGER*8, ALLOCATABLE :: off(:)
    REAL*8, ALLOCATABLE :: verts(:), nearest(:,:,:), normal(:,:), point(:,:), depth(:)
    INTEGER*4, ALLOCATABLE :: pa(:), pb(:), ctype(:)
    INTEGER*1, ALLOCATABLE :: col(:)
    INTEGER*4 :: version
    LOGICAL*1 :: flag
    CALL GET_COMMAND_ARGUMENT(1, fin); CALL GET_COMMAND_ARGUMENT(2, fout)
    version = 1
    OPEN(10, FILE=TRIM(fin), ACCESS="STREAM", FORM="UNFORMATTED", STATUS="OLD")
    READ(10) nh, nv, np
    ALLOCATE(off(nh + 1), verts(3 * nv), pa(np), pb(np))
    READ(10) off, verts, pa, pb
    CLOSE(10)
    ALLOCATE(col(np), ctype(np), nearest(2, 3, np), normal(3, np), point(3, np), depth(np))
    !$OMP PARALLEL DO SCHEDULE(DYNAMIC, 1024) PRIVATE(flag)
    DO i = 1, np
        CALL one(i)
    END DO
    !$OMP END PARALLEL DO
    OPEN(11, FILE=TRIM(fout), ACCESS="STREAM", FORM="UNFORMATTED", STATUS="REPLACE")
    WRITE(11) col, ctype, nearest, normal, point, depth
    CLOSE(11)
CONTAINS
    SUBROUTINE one(k)
        INTEGER*8, INTENT(IN) :: k
        INTEGER*8 :: a, b, na, nb
        REAL*8, ALLOCATABLE :: p1(:,:), p2(:,:)
        LOGICAL*1 :: f
        a = pa(k) + 1; b = pb(k) + 1
        na = off(a + 1) - off(a); nb = off(b + 1) - off(b)
        ! a hull block [x(n)|y(n)|z(n)] IS the column-major p(n,3)
        p1 = RESHAPE(verts(3 * off(a) + 1 : 3 * off(a) + 3 * na), [na, 3_8])
        p2 = RESHAPE(verts(3 * off(b) + 1 : 3 * off(b) + 3 * nb), [nb, 3_8])
        CALL GJKEPA(version, 1.D0, p1, p2, f, ctype(k), nearest(:, :, k), normal(:, k), point(:, k), depth(k))
        col(k) = MERGE(1_1, 0_1, f)
    END SUBROUTINE
END PROGRAM ref_driver
