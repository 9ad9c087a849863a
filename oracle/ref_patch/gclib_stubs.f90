! Stand-ins for the three modules GCLIB_GJKEPA USEs but its repository does not contain (F:13-15).
! UNTESTED here (no Fortran compiler in the image); part of the ref_patch recipe, not of the product.

MODULE GCLIB_List
    IMPLICIT NONE
    PRIVATE
    PUBLIC :: List_Array3d, ListNode_Array3d
    TYPE ListNode_Array3d
        REAL*8 :: value(3) = 0.D0
    END TYPE
    ! the reference uses the list as an append-only array: reset / append / getSize / atIndex (F:463-688)
    TYPE List_Array3d
        INTEGER*4 :: n = 0
        REAL*8, ALLOCATABLE :: v(:,:)
    CONTAINS
        PROCEDURE :: reset => list_reset
        PROCEDURE :: append => list_append
        PROCEDURE :: getSize => list_size
        PROCEDURE :: atIndex => list_at
    END TYPE
CONTAINS
    SUBROUTINE list_reset(this)
        CLASS(List_Array3d), INTENT(INOUT) :: this
        this%n = 0
    END SUBROUTINE
    SUBROUTINE list_append(this, p)
        CLASS(List_Array3d), INTENT(INOUT) :: this
        REAL*8, INTENT(IN) :: p(3)
        REAL*8, ALLOCATABLE :: t(:,:)
        IF (.NOT. ALLOCATED(this%v)) ALLOCATE(this%v(3, 16))
        IF (this%n == SIZE(this%v, 2)) THEN
            ALLOCATE(t(3, 2 * this%n)); t(:, 1:this%n) = this%v; CALL MOVE_ALLOC(t, this%v)
        END IF
        this%n = this%n + 1
        this%v(:, this%n) = p
    END SUBROUTINE
    INTEGER*4 FUNCTION list_size(this)
        CLASS(List_Array3d), INTENT(IN) :: this
        list_size = this%n
    END FUNCTION
    FUNCTION list_at(this, i) RESULT(node)
        CLASS(List_Array3d), INTENT(IN) :: this
        INTEGER*4, INTENT(IN) :: i
        TYPE(ListNode_Array3d) :: node
        node%value = this%v(:, i)
    END FUNCTION
END MODULE GCLIB_List

MODULE GCLIB_QuickHull
    USE, INTRINSIC :: ISO_C_BINDING
    IMPLICIT NONE
    PRIVATE
    PUBLIC :: QuickHull
    INTERFACE   ! oracle/gjkepa_oracle.h: pts (npts,3) row-major -> tris (ntri,9) row-major; 0 on success
        FUNCTION orc_quickhull(pts, npts, tris, ntri, cap) BIND(C, NAME="orc_quickhull") RESULT(rc)
            IMPORT :: C_DOUBLE, C_INT32_T, C_INT
            REAL(C_DOUBLE), INTENT(IN) :: pts(*)
            INTEGER(C_INT32_T), VALUE :: npts, cap
            REAL(C_DOUBLE), INTENT(OUT) :: tris(*)
            INTEGER(C_INT32_T), INTENT(OUT) :: ntri
            INTEGER(C_INT) :: rc
        END FUNCTION
    END INTERFACE
CONTAINS
    ! call site F:950: CALL QuickHull(scatPoints(n,3), polytope_2_(f,3,3) allocatable, info)
    SUBROUTINE QuickHull(points, polytope, info)
        REAL*8, INTENT(IN) :: points(:,:)
        REAL*8, INTENT(OUT), ALLOCATABLE :: polytope(:,:,:)
        INTEGER*4, INTENT(OUT) :: info
        INTEGER(C_INT32_T) :: n, nt, cap
        REAL(C_DOUBLE), ALLOCATABLE :: pts(:), tris(:)
        INTEGER :: i, f, v, c
        n = SIZE(points, 1); cap = 4 * n + 16
        ALLOCATE(pts(3 * n), tris(9 * cap))
        DO i = 1, n
            pts(3 * i - 2 : 3 * i) = points(i, :)
        END DO
        info = orc_quickhull(pts, n, tris, nt, cap)
        IF (info /= 0) nt = 0
        ALLOCATE(polytope(nt, 3, 3))
        DO f = 1, nt
            DO v = 1, 3
                DO c = 1, 3
                    polytope(f, v, c) = tris(9 * (f - 1) + 3 * (v - 1) + c)
                END DO
            END DO
        END DO
    END SUBROUTINE
END MODULE GCLIB_QuickHull

MODULE GCLIB_DeHull
    IMPLICIT NONE
    PRIVATE
    PUBLIC :: getHullMeshesVertex, getHullMeshesVertex_post
CONTAINS
    ! call site F:920: unique vertices of polytope(f,3,3), first appearance (face by face, vertex by vertex), exact compare
    SUBROUTINE getHullMeshesVertex(polytope, scatPoints, info)
        REAL*8, INTENT(IN) :: polytope(:,:,:)
        REAL*8, INTENT(OUT), ALLOCATABLE :: scatPoints(:,:)
        INTEGER*4, INTENT(OUT) :: info
        REAL*8, ALLOCATABLE :: t(:,:)
        INTEGER :: f, v, k, n
        LOGICAL :: seen
        ALLOCATE(t(3 * SIZE(polytope, 1), 3)); n = 0
        DO f = 1, SIZE(polytope, 1)
            DO v = 1, 3
                seen = .FALSE.
                DO k = 1, n
                    IF (ALL(t(k, :) == polytope(f, v, :))) THEN
                        seen = .TRUE.; EXIT
                    END IF
                END DO
                IF (.NOT. seen) THEN
                    n = n + 1; t(n, :) = polytope(f, v, :)
                END IF
            END DO
        END DO
        ALLOCATE(scatPoints(n, 3)); scatPoints = t(1:n, :)
        info = 0
    END SUBROUTINE
    SUBROUTINE getHullMeshesVertex_post()      ! imported by the reference, never called
    END SUBROUTINE
END MODULE GCLIB_DeHull
