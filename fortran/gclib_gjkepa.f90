!  gclib_gjkepa.f90 -- drop-in replacement for MODULE GCLIB_GJKEPA of
!  xiejihong0306/collision-detect-GJK-EPA (src/GCLIB_GJKEPA.f90), backed by libgjkepa_b200.so.
!
!  Host code keeps  USE GCLIB_GJKEPA  and  CALL GJKEPA(version_, TOL_FF_, p1_, p2_, collision_,
!  colliType_, nearest_points_, collision_normal_, collision_point_, penetration_depth_)  with the
!  dummy-argument list of the reference (lines 39-52 there); the body is a thin ISO_C_BINDING call
!  into the C ABI of include/gjkepa.h.  GJKEPA_BATCH is new: n_pairs independent pairs per call, which
!  is where the GPU pays off (one pair per call costs two PCIe round trips).
!
!  NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Fortran compiler here (see INTEGRATION.md);
!  tests/test_ref_fortran.py compiles it wherever gfortran is on PATH.  Build with e.g.
!  gfortran -c gclib_gjkepa.f90  and link  -L<repo>/collision_detect_gjk_epa_b200 -lgjkepa_b200.
!  The C ABI underneath is exercised by the test-suite through ctypes with exactly these conventions
!  (every scalar by reference, column-major arrays).
!
!  ERRORS.  Nothing in this module stops the program (the header's contract: "never exit").  Every
!  routine but GJKEPA takes an OPTIONAL last argument ierr that receives the library's return code
!  (0 = ok, 1 = bad argument, 2 = CUDA, 3 = out of device memory); without it a failure prints one
!  line on unit 0 and the outputs keep the reference's "no collision" values (F:65-71).  GJKEPA keeps
!  the reference's dummy list: on failure it returns collision_ = .FALSE. with zeroed outputs and
!  GJKEPA_LAST_STATUS() returns the code (per thread).
MODULE GCLIB_GJKEPA
    USE, INTRINSIC :: ISO_C_BINDING
    IMPLICIT NONE
    PRIVATE
    PUBLIC :: GJKEPA, GJKEPA_BATCH, GJKEPA_CTX_CREATE, GJKEPA_CTX_DESTROY, GJKEPA_SHAPES_UPLOAD, GJKEPA_BATCH_POSED, &
              GJKEPA_BROADPHASE, GJKEPA_BROADPHASE_POSED, GJKEPA_BATCH_R4, GJKEPA_SCENE_POSED, &
              GJKEPA_CTX_CREATE_MULTI, GJKEPA_STORE_UPLOAD, GJKEPA_BATCH_RESIDENT, GJKEPA_CONTACTS_RESIDENT, &
              GJKEPA_PIN, GJKEPA_UNPIN, GJKEPA_LAST_STATUS, GJKEPA_SET_TUNING, GJKEPA_SET_MAX_HULL_VERTICES

    INTEGER(C_INT32_T), PARAMETER, PUBLIC :: GJKEPA_MODE_GJK = 0, GJKEPA_MODE_EPA = 1, GJKEPA_MODE_FULL = 2
    INTEGER(C_INT32_T), PARAMETER, PUBLIC :: GJKEPA_MODE_COMPACT = 16       ! OR-ed into mode: one row per contact
    ! keys of GJKEPA_SET_TUNING (include/gjkepa.h GJKEPA_TUNE_*): none changes a result except as gjkepa.h states
    INTEGER(C_INT32_T), PARAMETER, PUBLIC :: GJKEPA_TUNE_EARLY_OUT = 5, GJKEPA_TUNE_HOST_CHUNKS = 6, &
                                             GJKEPA_TUNE_LOCALITY_SORT = 7, GJKEPA_TUNE_F32_FILTER = 8, GJKEPA_TUNE_F32_ARITH = 9

    INTEGER(C_INT), SAVE :: last_status = 0
    !$OMP THREADPRIVATE(last_status)

    INTERFACE
        FUNCTION gjkepa_pair_f64(version, tol_ff, p1, n1, p2, n2, collision, colli_type, &
                                 nearest, normal, point, depth) BIND(C, NAME="gjkepa_pair_f64") RESULT(rc)
            IMPORT :: C_INT, C_INT32_T, C_DOUBLE, C_INT8_T
            INTEGER(C_INT32_T), INTENT(IN)  :: version, n1, n2
            REAL(C_DOUBLE),     INTENT(IN)  :: tol_ff
            REAL(C_DOUBLE),     INTENT(IN)  :: p1(*), p2(*)          ! p(n,3) column-major = [x|y|z]
            INTEGER(C_INT8_T),  INTENT(OUT) :: collision
            INTEGER(C_INT32_T), INTENT(OUT) :: colli_type
            REAL(C_DOUBLE),     INTENT(OUT) :: nearest(2,3), normal(3), point(3), depth
            INTEGER(C_INT) :: rc
        END FUNCTION gjkepa_pair_f64

        FUNCTION gjkepa_create(ctx, device) BIND(C, NAME="gjkepa_create") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T
            TYPE(C_PTR), INTENT(OUT) :: ctx
            INTEGER(C_INT32_T), VALUE :: device
            INTEGER(C_INT) :: rc
        END FUNCTION gjkepa_create

        FUNCTION gjkepa_create_multi(ctx, devices, n_devices) BIND(C, NAME="gjkepa_create_multi") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T
            TYPE(C_PTR), INTENT(OUT) :: ctx
            TYPE(C_PTR), VALUE :: devices
            INTEGER(C_INT32_T), VALUE :: n_devices
            INTEGER(C_INT) :: rc
        END FUNCTION gjkepa_create_multi
        FUNCTION gjkepa_set_tuning(ctx, key, val) BIND(C, NAME="gjkepa_set_tuning") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T
            TYPE(C_PTR), VALUE :: ctx
            INTEGER(C_INT32_T), VALUE :: key, val
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_set_max_hull_vertices(ctx, n) BIND(C, NAME="gjkepa_set_max_hull_vertices") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T
            TYPE(C_PTR), VALUE :: ctx
            INTEGER(C_INT32_T), VALUE :: n
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_store_upload(ctx, n_hulls, verts, hull_off) BIND(C, NAME="gjkepa_store_upload") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT64_T
            TYPE(C_PTR), VALUE :: ctx, verts, hull_off
            INTEGER(C_INT64_T), VALUE :: n_hulls
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_batch_resident_f64(ctx, version, tol_ff, n_pairs, pair_a, pair_b, mode, collision, colli_type, nearest, &
                                           normal, point, depth, status, n_contacts, contact_index) &
                                           BIND(C, NAME="gjkepa_batch_resident_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
            TYPE(C_PTR), VALUE :: ctx, pair_a, pair_b
            TYPE(C_PTR), VALUE :: collision, colli_type, nearest, normal, point, depth, status, n_contacts, contact_index
            INTEGER(C_INT32_T), VALUE :: version, mode
            REAL(C_DOUBLE), VALUE :: tol_ff
            INTEGER(C_INT64_T), VALUE :: n_pairs
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_host_register(p, bytes) BIND(C, NAME="gjkepa_host_register") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT64_T
            TYPE(C_PTR), VALUE :: p
            INTEGER(C_INT64_T), VALUE :: bytes
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_host_unregister(p) BIND(C, NAME="gjkepa_host_unregister") RESULT(rc)
            IMPORT :: C_PTR, C_INT
            TYPE(C_PTR), VALUE :: p
            INTEGER(C_INT) :: rc
        END FUNCTION

        FUNCTION gjkepa_destroy(ctx) BIND(C, NAME="gjkepa_destroy") RESULT(rc)
            IMPORT :: C_PTR, C_INT
            TYPE(C_PTR), VALUE :: ctx
            INTEGER(C_INT) :: rc
        END FUNCTION gjkepa_destroy

        FUNCTION gjkepa_batch_f64(ctx, version, tol_ff, n_hulls, verts, hull_off, n_pairs, pair_a, pair_b, &
                                  mode, collision, colli_type, nearest, normal, point, depth, status, &
                                  n_contacts, contact_index) BIND(C, NAME="gjkepa_batch_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
            TYPE(C_PTR), VALUE :: ctx
            INTEGER(C_INT32_T), VALUE :: version, mode
            REAL(C_DOUBLE), VALUE :: tol_ff
            INTEGER(C_INT64_T), VALUE :: n_hulls, n_pairs
            TYPE(C_PTR), VALUE :: verts, hull_off, pair_a, pair_b
            TYPE(C_PTR), VALUE :: collision, colli_type, nearest, normal, point, depth, status
            TYPE(C_PTR), VALUE :: n_contacts, contact_index
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_shapes_upload(ctx, n_shapes, shape_verts, shape_off) BIND(C, NAME="gjkepa_shapes_upload") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT64_T
            TYPE(C_PTR), VALUE :: ctx, shape_verts, shape_off
            INTEGER(C_INT64_T), VALUE :: n_shapes
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_batch_posed_f64(ctx, version, tol_ff, n_bodies, body_shape, pose, n_pairs, pair_a, pair_b, mode, &
                                        collision, colli_type, nearest, normal, point, depth, status, &
                                        n_contacts, contact_index) BIND(C, NAME="gjkepa_batch_posed_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
            TYPE(C_PTR), VALUE :: ctx, body_shape, pose, pair_a, pair_b
            TYPE(C_PTR), VALUE :: collision, colli_type, nearest, normal, point, depth, status, n_contacts, contact_index
            INTEGER(C_INT32_T), VALUE :: version, mode
            REAL(C_DOUBLE), VALUE :: tol_ff
            INTEGER(C_INT64_T), VALUE :: n_bodies, n_pairs
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_batch_f32(ctx, version, tol_ff, n_hulls, verts, hull_off, n_pairs, pair_a, pair_b, mode, &
                                  collision, colli_type, nearest, normal, point, depth, status, &
                                  n_contacts, contact_index) BIND(C, NAME="gjkepa_batch_f32") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_FLOAT
            TYPE(C_PTR), VALUE :: ctx, verts, hull_off, pair_a, pair_b
            TYPE(C_PTR), VALUE :: collision, colli_type, nearest, normal, point, depth, status, n_contacts, contact_index
            INTEGER(C_INT32_T), VALUE :: version, mode
            REAL(C_FLOAT), VALUE :: tol_ff
            INTEGER(C_INT64_T), VALUE :: n_hulls, n_pairs
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_scene_posed_f64(ctx, version, tol_ff, n_bodies, body_shape, pose, mode, cap, body_a, body_b, depth, &
                                        normal, point, nearest, colli_type, n_candidates, n_contacts) &
                                        BIND(C, NAME="gjkepa_scene_posed_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
            TYPE(C_PTR), VALUE :: ctx, body_shape, pose, body_a, body_b, depth, normal, point, nearest, colli_type
            INTEGER(C_INT32_T), VALUE :: version, mode
            REAL(C_DOUBLE), VALUE :: tol_ff
            INTEGER(C_INT64_T), VALUE :: n_bodies, cap
            INTEGER(C_INT64_T) :: n_candidates, n_contacts
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_broadphase_f64(ctx, n_hulls, verts, hull_off, cap, pair_a, pair_b, n_pairs) &
                                       BIND(C, NAME="gjkepa_broadphase_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT64_T
            TYPE(C_PTR), VALUE :: ctx, verts, hull_off, pair_a, pair_b
            INTEGER(C_INT64_T), VALUE :: n_hulls, cap
            INTEGER(C_INT64_T) :: n_pairs
            INTEGER(C_INT) :: rc
        END FUNCTION
        FUNCTION gjkepa_broadphase_posed_f64(ctx, n_bodies, body_shape, pose, cap, pair_a, pair_b, n_pairs) &
                                             BIND(C, NAME="gjkepa_broadphase_posed_f64") RESULT(rc)
            IMPORT :: C_PTR, C_INT, C_INT64_T
            TYPE(C_PTR), VALUE :: ctx, body_shape, pose, pair_a, pair_b
            INTEGER(C_INT64_T), VALUE :: n_bodies, cap
            INTEGER(C_INT64_T) :: n_pairs
            INTEGER(C_INT) :: rc
        END FUNCTION
    END INTERFACE

CONTAINS

    ! hands the library's return code to the caller (ierr present) or reports it on unit 0; never stops the program
    SUBROUTINE report(rc, ierr, what)
        INTEGER(C_INT), INTENT(IN) :: rc
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CHARACTER(LEN=*), INTENT(IN) :: what
        last_status = rc
        IF (PRESENT(ierr)) THEN
            ierr = INT(rc)
        ELSE IF (rc /= 0) THEN
            WRITE(0, '(A,A,A,I0,A)') "GCLIB_GJKEPA: ", what, " failed with code ", rc, &
                                     " (no CPU fallback; outputs left as no-collision)"
        END IF
    END SUBROUTINE report

    INTEGER FUNCTION GJKEPA_LAST_STATUS()
        GJKEPA_LAST_STATUS = INT(last_status)
    END FUNCTION GJKEPA_LAST_STATUS

    ! Same dummy list as the reference's GJKEPA.  p1_/p2_ stay assumed-shape; passing them to the
    ! explicit-size BIND(C) dummies makes the compiler hand over contiguous column-major copies.
    SUBROUTINE GJKEPA(version_, TOL_FF_, p1_, p2_, collision_, colliType_, &
                      nearest_points_, collision_normal_, collision_point_, penetration_depth_)
        INTEGER*4, INTENT(IN)  :: version_
        REAL*8,    INTENT(IN)  :: TOL_FF_
        REAL*8,    INTENT(IN)  :: p1_(:,:), p2_(:,:)
        LOGICAL*1, INTENT(OUT) :: collision_
        INTEGER*4, INTENT(OUT) :: colliType_
        REAL*8,    INTENT(OUT) :: nearest_points_(2,3), collision_normal_(3), collision_point_(3)
        REAL*8,    INTENT(OUT) :: penetration_depth_
        INTEGER(C_INT8_T)  :: flag
        INTEGER(C_INT32_T) :: n1, n2
        INTEGER(C_INT)     :: rc
        n1 = INT(SIZE(p1_, 1), C_INT32_T)
        n2 = INT(SIZE(p2_, 1), C_INT32_T)
        rc = gjkepa_pair_f64(version_, TOL_FF_, p1_, n1, p2_, n2, flag, colliType_, &
                             nearest_points_, collision_normal_, collision_point_, penetration_depth_)
        last_status = rc
        IF (rc /= 0) THEN               ! the reference's own failure exits all answer .FALSE. (quirk Q5); outputs as F:65-71
            collision_ = .FALSE.; colliType_ = 0; nearest_points_ = 0.D0; collision_normal_ = 0.D0
            collision_point_ = 0.D0; penetration_depth_ = 0.D0
            RETURN
        END IF
        collision_ = (flag /= 0)        ! the library writes 1; Intel's .TRUE. is -1, so test "/= 0"
    END SUBROUTINE GJKEPA

    SUBROUTINE GJKEPA_CTX_CREATE(ctx, device, ierr)
        TYPE(C_PTR), INTENT(OUT) :: ctx
        INTEGER*4, INTENT(IN) :: device
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_create(ctx, INT(device, C_INT32_T)), ierr, "GJKEPA_CTX_CREATE")
    END SUBROUTINE GJKEPA_CTX_CREATE

    ! Knobs of a context (keys GJKEPA_TUNE_* above): kernel selection, host pipeline, the opt-in shortcut E1, the fp32 entry's
    ! binary32 filter / binary32 walk.  A multi-device context passes them on to every device.
    SUBROUTINE GJKEPA_SET_TUNING(ctx, key, val, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER*4, INTENT(IN) :: key, val
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_set_tuning(ctx, INT(key, C_INT32_T), INT(val, C_INT32_T)), ierr, "GJKEPA_SET_TUNING")
    END SUBROUTINE GJKEPA_SET_TUNING

    ! The caller's bound on vertices per hull (0 = unknown): sizes the shared-memory columns.  A pair that violates it is
    ! reported in its status word (GJKEPA_ST_BAD_HINT, flag .FALSE.), never processed silently.
    SUBROUTINE GJKEPA_SET_MAX_HULL_VERTICES(ctx, n, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER*4, INTENT(IN) :: n
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_set_max_hull_vertices(ctx, INT(n, C_INT32_T)), ierr, "GJKEPA_SET_MAX_HULL_VERTICES")
    END SUBROUTINE GJKEPA_SET_MAX_HULL_VERTICES

    ! One context over several GPUs of the box (devices absent: every visible GPU).  The batch routines below then split
    ! the pair list into one contiguous range per GPU inside the library and every GPU copies its results straight into
    ! the caller's arrays: the reference's "!$OMP PARALLEL DO over pairs" (F:55-60) with GPUs in place of cores.
    SUBROUTINE GJKEPA_CTX_CREATE_MULTI(ctx, devices, ierr)
        TYPE(C_PTR), INTENT(OUT) :: ctx
        INTEGER(C_INT32_T), INTENT(IN), OPTIONAL, TARGET, CONTIGUOUS :: devices(:)
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        IF (PRESENT(devices)) THEN
            CALL report(gjkepa_create_multi(ctx, C_LOC(devices), INT(SIZE(devices), C_INT32_T)), ierr, "GJKEPA_CTX_CREATE_MULTI")
        ELSE
            CALL report(gjkepa_create_multi(ctx, C_NULL_PTR, 0_C_INT32_T), ierr, "GJKEPA_CTX_CREATE_MULTI")
        END IF
    END SUBROUTINE GJKEPA_CTX_CREATE_MULTI

    ! Page-lock an array that is passed to the batch routines step after step (ALLOCATEd arrays are pageable: their copies
    ! run at a fraction of PCIe speed and cannot overlap the kernels).  GJKEPA_UNPIN before DEALLOCATE.
    SUBROUTINE GJKEPA_PIN(addr, bytes, ierr)
        TYPE(C_PTR), INTENT(IN) :: addr              ! C_LOC(array)
        INTEGER(C_INT64_T), INTENT(IN) :: bytes
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_host_register(addr, bytes), ierr, "GJKEPA_PIN")
    END SUBROUTINE GJKEPA_PIN
    SUBROUTINE GJKEPA_UNPIN(addr, ierr)
        TYPE(C_PTR), INTENT(IN) :: addr
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_host_unregister(addr), ierr, "GJKEPA_UNPIN")
    END SUBROUTINE GJKEPA_UNPIN

    SUBROUTINE GJKEPA_CTX_DESTROY(ctx)
        TYPE(C_PTR), INTENT(INOUT) :: ctx
        INTEGER(C_INT) :: rc
        rc = gjkepa_destroy(ctx)
        ctx = C_NULL_PTR
    END SUBROUTINE GJKEPA_CTX_DESTROY

    ! verts: hull blocks back to back, hull h = p(n_h,3) column-major starting at 3*hull_off(h)+1
    ! (hull_off is 0-based, in vertices, n_hulls+1 entries).  pair_a/pair_b: 0-based hull numbers.
    SUBROUTINE GJKEPA_BATCH(ctx, version_, TOL_FF_, verts, hull_off, pair_a, pair_b, mode, &
                            collision_, colliType_, nearest_points_, collision_normal_, &
                            collision_point_, penetration_depth_, status_, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER*4, INTENT(IN) :: version_, mode
        REAL*8,    INTENT(IN) :: TOL_FF_
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: verts(:)
        INTEGER(C_INT64_T), INTENT(IN), TARGET, CONTIGUOUS :: hull_off(:)
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT8_T),  INTENT(OUT), TARGET, CONTIGUOUS :: collision_(:)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: colliType_(:), status_(:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: nearest_points_(:,:,:)   ! (2,3,n_pairs)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: collision_normal_(:,:)   ! (3,n_pairs)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: collision_point_(:,:)    ! (3,n_pairs)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: penetration_depth_(:)
        INTEGER(C_INT) :: rc
        rc = gjkepa_batch_f64(ctx, version_, TOL_FF_, INT(SIZE(hull_off) - 1, C_INT64_T), C_LOC(verts), &
                              C_LOC(hull_off), INT(SIZE(pair_a), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), mode, &
                              C_LOC(collision_), C_LOC(colliType_), C_LOC(nearest_points_), &
                              C_LOC(collision_normal_), C_LOC(collision_point_), C_LOC(penetration_depth_), &
                              C_LOC(status_), C_NULL_PTR, C_NULL_PTR)
        CALL report(rc, ierr, "GJKEPA_BATCH")
    END SUBROUTINE GJKEPA_BATCH

    ! Shapes stay on the GPU, only poses cross PCIe per call.  shape_verts/shape_off: a hull store of BODY-space
    ! shapes (layout as verts/hull_off above), copied to the device once.
    SUBROUTINE GJKEPA_SHAPES_UPLOAD(ctx, shape_verts, shape_off, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: shape_verts(:)
        INTEGER(C_INT64_T), INTENT(IN), TARGET, CONTIGUOUS :: shape_off(:)
        INTEGER(C_INT) :: rc
        rc = gjkepa_shapes_upload(ctx, INT(SIZE(shape_off) - 1, C_INT64_T), C_LOC(shape_verts), C_LOC(shape_off))
        CALL report(rc, ierr, "GJKEPA_SHAPES_UPLOAD")
    END SUBROUTINE GJKEPA_SHAPES_UPLOAD

    ! body_shape(b): 0-based shape number of body b; pose(:,b) = r00 r01 r02 r10 r11 r12 r20 r21 r22 tx ty tz
    ! (R row-major; world = ((r00*x + r01*y) + r02*z) + tx, applied on the device); pair_a/pair_b: 0-based BODY numbers.
    SUBROUTINE GJKEPA_BATCH_POSED(ctx, version_, TOL_FF_, body_shape, pose, pair_a, pair_b, mode, &
                                  collision_, colliType_, nearest_points_, collision_normal_, &
                                  collision_point_, penetration_depth_, status_, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER*4, INTENT(IN) :: version_, mode
        REAL*8,    INTENT(IN) :: TOL_FF_
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: body_shape(:), pair_a(:), pair_b(:)
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: pose(:,:)                  ! (12,n_bodies)
        INTEGER(C_INT8_T),  INTENT(OUT), TARGET, CONTIGUOUS :: collision_(:)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: colliType_(:), status_(:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: nearest_points_(:,:,:), collision_normal_(:,:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: collision_point_(:,:), penetration_depth_(:)
        INTEGER(C_INT) :: rc
        rc = gjkepa_batch_posed_f64(ctx, version_, TOL_FF_, INT(SIZE(body_shape), C_INT64_T), C_LOC(body_shape), &
                                    C_LOC(pose), INT(SIZE(pair_a), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), mode, &
                                    C_LOC(collision_), C_LOC(colliType_), C_LOC(nearest_points_), &
                                    C_LOC(collision_normal_), C_LOC(collision_point_), C_LOC(penetration_depth_), &
                                    C_LOC(status_), C_NULL_PTR, C_NULL_PTR)
        CALL report(rc, ierr, "GJKEPA_BATCH_POSED")
    END SUBROUTINE GJKEPA_BATCH_POSED

    ! Candidate pairs on the device: every pair of bodies (a < b, 0-based) that passes the reference's own rough test
    ! (RoughCollisionDetection_SphericalEnvelope: NORM2(mp1 - mp2) <= r1 + r2 + 1.0), found through a uniform grid
    ! instead of the host's loop over all pairs.  n_pairs may exceed SIZE(pair_a): then only the first SIZE(pair_a)
    ! were written and the call is repeated with larger arrays.
    SUBROUTINE GJKEPA_BROADPHASE(ctx, verts, hull_off, pair_a, pair_b, n_pairs, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: verts(:)
        INTEGER(C_INT64_T), INTENT(IN), TARGET, CONTIGUOUS :: hull_off(:)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT64_T), INTENT(OUT) :: n_pairs
        INTEGER(C_INT) :: rc
        rc = gjkepa_broadphase_f64(ctx, INT(SIZE(hull_off) - 1, C_INT64_T), C_LOC(verts), C_LOC(hull_off), &
                                   INT(MIN(SIZE(pair_a), SIZE(pair_b)), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), n_pairs)
        CALL report(rc, ierr, "GJKEPA_BROADPHASE")
    END SUBROUTINE GJKEPA_BROADPHASE

    SUBROUTINE GJKEPA_BROADPHASE_POSED(ctx, body_shape, pose, pair_a, pair_b, n_pairs, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: body_shape(:)
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: pose(:,:)                  ! (12,n_bodies)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT64_T), INTENT(OUT) :: n_pairs
        INTEGER(C_INT) :: rc
        rc = gjkepa_broadphase_posed_f64(ctx, INT(SIZE(body_shape), C_INT64_T), C_LOC(body_shape), C_LOC(pose), &
                                         INT(MIN(SIZE(pair_a), SIZE(pair_b)), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), n_pairs)
        CALL report(rc, ierr, "GJKEPA_BROADPHASE_POSED")
    END SUBROUTINE GJKEPA_BROADPHASE_POSED

    ! fp32 mode for hosts that keep REAL*4 geometry: float hull store and float results cross PCIe, the device widens
    ! the store and runs the same REAL*8 kernels (flags = GJKEPA's on the float-rounded vertices; reals within 1e-5).
    SUBROUTINE GJKEPA_BATCH_R4(ctx, version_, TOL_FF_, verts, hull_off, pair_a, pair_b, mode, &
                               collision_, colliType_, nearest_points_, collision_normal_, &
                               collision_point_, penetration_depth_, status_, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER*4, INTENT(IN) :: version_, mode
        REAL*4,    INTENT(IN) :: TOL_FF_
        REAL(C_FLOAT),      INTENT(IN), TARGET, CONTIGUOUS :: verts(:)
        INTEGER(C_INT64_T), INTENT(IN), TARGET, CONTIGUOUS :: hull_off(:)
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT8_T),  INTENT(OUT), TARGET, CONTIGUOUS :: collision_(:)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: colliType_(:), status_(:)
        REAL(C_FLOAT),      INTENT(OUT), TARGET, CONTIGUOUS :: nearest_points_(:,:,:), collision_normal_(:,:)
        REAL(C_FLOAT),      INTENT(OUT), TARGET, CONTIGUOUS :: collision_point_(:,:), penetration_depth_(:)
        INTEGER(C_INT) :: rc
        rc = gjkepa_batch_f32(ctx, version_, TOL_FF_, INT(SIZE(hull_off) - 1, C_INT64_T), C_LOC(verts), C_LOC(hull_off), &
                              INT(SIZE(pair_a), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), mode, &
                              C_LOC(collision_), C_LOC(colliType_), C_LOC(nearest_points_), &
                              C_LOC(collision_normal_), C_LOC(collision_point_), C_LOC(penetration_depth_), &
                              C_LOC(status_), C_NULL_PTR, C_NULL_PTR)
        CALL report(rc, ierr, "GJKEPA_BATCH_R4")
    END SUBROUTINE GJKEPA_BATCH_R4

    ! One simulation step in one call: poses in (shapes resident), CONTACTS out -- pose application, candidate pairs by the
    ! reference's sphere test and GJKEPA over them all happen on the GPU.  Rows 1..MIN(n_contacts, SIZE(body_a)) are written:
    ! body_a/body_b (0-based, a < b) and that pair's penetration_depth_ / collision_normal_ (mode >= GJKEPA_MODE_EPA).
    SUBROUTINE GJKEPA_SCENE_POSED(ctx, version_, TOL_FF_, body_shape, pose, mode, body_a, body_b, &
                                  penetration_depth_, collision_normal_, n_candidates, n_contacts, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER*4, INTENT(IN) :: version_, mode
        REAL*8,    INTENT(IN) :: TOL_FF_
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: body_shape(:)
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: pose(:,:)                  ! (12,n_bodies)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: body_a(:), body_b(:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: penetration_depth_(:), collision_normal_(:,:)   ! (3,cap)
        INTEGER(C_INT64_T), INTENT(OUT) :: n_candidates, n_contacts
        INTEGER(C_INT) :: rc
        rc = gjkepa_scene_posed_f64(ctx, version_, TOL_FF_, INT(SIZE(body_shape), C_INT64_T), C_LOC(body_shape), &
                                    C_LOC(pose), mode, &
                                    INT(MIN(SIZE(body_a), SIZE(body_b), SIZE(penetration_depth_)), C_INT64_T), C_LOC(body_a), &
                                    C_LOC(body_b), C_LOC(penetration_depth_), C_LOC(collision_normal_), C_NULL_PTR, C_NULL_PTR, &
                                    C_NULL_PTR, n_candidates, n_contacts)
        CALL report(rc, ierr, "GJKEPA_SCENE_POSED")
    END SUBROUTINE GJKEPA_SCENE_POSED

    ! World-space hull store resident on the GPU(s) of the context (copied once; replicated on a multi-GPU context, every
    ! GPU over its own PCIe link): afterwards only pair lists go in and results come out.
    SUBROUTINE GJKEPA_STORE_UPLOAD(ctx, verts, hull_off, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        REAL(C_DOUBLE),     INTENT(IN), TARGET, CONTIGUOUS :: verts(:)
        INTEGER(C_INT64_T), INTENT(IN), TARGET, CONTIGUOUS :: hull_off(:)
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        CALL report(gjkepa_store_upload(ctx, INT(SIZE(hull_off) - 1, C_INT64_T), C_LOC(verts), C_LOC(hull_off)), &
                    ierr, "GJKEPA_STORE_UPLOAD")
    END SUBROUTINE GJKEPA_STORE_UPLOAD

    ! GJKEPA_BATCH against the resident store: results at the pair's index, zero for non-colliding pairs (F:65-71).
    SUBROUTINE GJKEPA_BATCH_RESIDENT(ctx, version_, TOL_FF_, pair_a, pair_b, mode, collision_, colliType_, nearest_points_, &
                                     collision_normal_, collision_point_, penetration_depth_, status_, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER*4, INTENT(IN) :: version_, mode
        REAL*8,    INTENT(IN) :: TOL_FF_
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT8_T),  INTENT(OUT), TARGET, CONTIGUOUS :: collision_(:)
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: colliType_(:), status_(:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: nearest_points_(:,:,:), collision_normal_(:,:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: collision_point_(:,:), penetration_depth_(:)
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER(C_INT) :: rc
        rc = gjkepa_batch_resident_f64(ctx, version_, TOL_FF_, INT(SIZE(pair_a), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), mode, &
                                       C_LOC(collision_), C_LOC(colliType_), C_LOC(nearest_points_), C_LOC(collision_normal_), &
                                       C_LOC(collision_point_), C_LOC(penetration_depth_), C_LOC(status_), C_NULL_PTR, C_NULL_PTR)
        CALL report(rc, ierr, "GJKEPA_BATCH_RESIDENT")
    END SUBROUTINE GJKEPA_BATCH_RESIDENT

    ! The contact list of a pair batch against the resident store: rows 1..n_contacts of contact_pair (0-based pair
    ! numbers, ascending), penetration_depth_ and collision_normal_ (GJKEPA's outputs for that pair); collision_ stays
    ! one byte per pair.  Only contacts cross PCIe.  All row arrays need SIZE(pair_a) rows at worst.
    SUBROUTINE GJKEPA_CONTACTS_RESIDENT(ctx, version_, TOL_FF_, pair_a, pair_b, collision_, n_contacts, contact_pair, &
                                        penetration_depth_, collision_normal_, ierr)
        TYPE(C_PTR), INTENT(IN) :: ctx
        INTEGER*4, INTENT(IN) :: version_
        REAL*8,    INTENT(IN) :: TOL_FF_
        INTEGER(C_INT32_T), INTENT(IN), TARGET, CONTIGUOUS :: pair_a(:), pair_b(:)
        INTEGER(C_INT8_T),  INTENT(OUT), TARGET, CONTIGUOUS :: collision_(:)
        INTEGER(C_INT64_T), INTENT(OUT), TARGET :: n_contacts
        INTEGER(C_INT32_T), INTENT(OUT), TARGET, CONTIGUOUS :: contact_pair(:)
        REAL(C_DOUBLE),     INTENT(OUT), TARGET, CONTIGUOUS :: penetration_depth_(:), collision_normal_(:,:)   ! (3,rows)
        INTEGER, INTENT(OUT), OPTIONAL :: ierr
        INTEGER(C_INT) :: rc
        rc = gjkepa_batch_resident_f64(ctx, version_, TOL_FF_, INT(SIZE(pair_a), C_INT64_T), C_LOC(pair_a), C_LOC(pair_b), &
                                       IOR(GJKEPA_MODE_EPA, GJKEPA_MODE_COMPACT), C_LOC(collision_), C_NULL_PTR, C_NULL_PTR, &
                                       C_LOC(collision_normal_), C_NULL_PTR, C_LOC(penetration_depth_), C_NULL_PTR, &
                                       C_LOC(n_contacts), C_LOC(contact_pair))
        CALL report(rc, ierr, "GJKEPA_CONTACTS_RESIDENT")
    END SUBROUTINE GJKEPA_CONTACTS_RESIDENT

END MODULE GCLIB_GJKEPA
