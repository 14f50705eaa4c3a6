/*
 * constv_b200.h -- C ABI of the B200-native ConstVLangevinIntegrator step path.
 *
 * This is the drop-in boundary under the reference's OpenMM plugin classes: the C++ kernel class
 * (openmm_constv_b200/plugin/platforms/cuda) and the Python host mirror (openmm_constv_b200/) both
 * call ONLY these entry points.  Plain pointers, sizes and scalars; no C++ or torch types.
 * Every function cites the reference interface it replaces; paths are relative to the reference
 * repository root (scychon/openmm_constV), "Kernels.cpp" = platforms/cuda/src/CudaConstVKernels.cpp,
 * "Langevin.cu" = platforms/cuda/src/kernels/constVLangevin.cu.
 *
 * Conventions
 *   - All functions return CONSTV_OK (0) or an error code; constv_last_error() gives the message of
 *     the calling thread's last failure.  A CONSTV_ERR_SYSTEM message is verbatim the text of the
 *     OpenMMException the reference throws for the same condition.
 *   - `stream` is a cudaStream_t / CUstream passed as void* (NULL = legacy default stream, which is
 *     what OpenMM 7.x launches on).  All launches are asynchronous unless stated.
 *   - Device arrays use OpenMM's CUDA layouts, P = padded_num_atoms:
 *       posq            real4[P]   x, y, z, charge
 *       posq_correction real4[P]   mixed precision only, else NULL
 *       velm            mixed4[P]  vx, vy, vz, inverse mass (0 = massless)
 *       force           int64[3*P] x-plane, y-plane, z-plane, fixed point scaled by 2^32
 *       pos_delta       mixed4[P]  only for the split (constraint) path
 *       atom_index      int32[P]   slot -> original atom index, NULL = identity order
 *     real/mixed = float/float (single), float/double (mixed), double/double (double).
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     CONSTV_ERR_CUDA.
 */
#ifndef CONSTV_B200_H_
#define CONSTV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CONSTV_API __attribute__((visibility("default")))

#define CONSTV_VERSION 121 /* 0.1.2: + constv_mirror, constv_generate_philox_words */

typedef struct constv_slab constv_slab; /* one Context's worth of integrator state (opaque) */

enum {
    CONSTV_OK = 0,
    CONSTV_ERR_INVALID = 1, /* bad argument */
    CONSTV_ERR_CUDA = 2,    /* CUDA runtime failure (message carries cudaGetErrorString) */
    CONSTV_ERR_STATE = 3,   /* call sequence error: arrays not bound, parameters not set, ... */
    CONSTV_ERR_SYSTEM = 4   /* the System violates the integrator's requirements (reference text) */
};

enum { CONSTV_PRECISION_SINGLE = 0, CONSTV_PRECISION_MIXED = 1, CONSTV_PRECISION_DOUBLE = 2 };

enum {
    /* z_img = -z + 2*i*zmax, evaluated in double, as the shipped kernel does (Langevin.cu:151-152) */
    CONSTV_MIRROR_REFERENCE = 0,
    /* z_img = -z exactly (BASELINE.json north_star); numCells must be 2 */
    CONSTV_MIRROR_NEGATE = 1
};

enum {
    CONSTV_FLAG_ZERO_IMAGE_VELM = 1u << 0,   /* rewrite every image's velm to (0,0,0,0) each step */
    CONSTV_FLAG_ZERO_IMAGE_FORCE = 1u << 1,  /* rewrite every image's three force entries to 0   */
    /* read image charges from a side array snapshotted by constv_refresh_image_charges() instead of
     * re-reading posq[image].w every step (Langevin.cu:153).  Charges are constant during a run
     * unless the caller edits them (then call constv_refresh_image_charges again). */
    CONSTV_FLAG_CACHE_IMAGE_CHARGE = 1u << 2,
    /* launch with programmatic dependent launch so back-to-back steps overlap launch latency */
    CONSTV_FLAG_PDL = 1u << 3,
    /* By default the step marks everything it touches evict-first (ld/st.global.cs): in a real
     * simulation the force kernels between two steps flush L2 anyway.  Set this for step-only loops
     * over a slab that fits in L2 (56 B per atom) so its state stays cache resident. */
    CONSTV_FLAG_KEEP_IN_L2 = 1u << 4
};
#define CONSTV_FLAGS_DEFAULT (CONSTV_FLAG_ZERO_IMAGE_VELM | CONSTV_FLAG_ZERO_IMAGE_FORCE | CONSTV_FLAG_PDL)

typedef struct constv_config {
    uint32_t struct_size;      /* = sizeof(constv_config), for ABI evolution */
    int32_t device;            /* CUDA device ordinal; -1 = current device */
    int32_t num_atoms;         /* N = all particles, real + image (cu.getNumAtoms()) */
    int32_t padded_num_atoms;  /* P >= N, force plane stride (cu.getPaddedNumAtoms()) */
    int32_t num_cells;         /* integrator.getNumCells(): even, N % num_cells == 0 */
    int32_t precision;         /* CONSTV_PRECISION_* */
    int32_t mirror_mode;       /* CONSTV_MIRROR_* */
    uint32_t flags;            /* CONSTV_FLAG_* */
    double zmax;               /* resolved cell size (> 0), see constv_validate_system */
    uint64_t seed;             /* Philox key; integrator.getRandomNumberSeed() (0 = caller picks) */
    uint64_t global_atom_offset; /* original index of this slab's atom 0 in the whole system
                                    (multi-GPU range partition); keys the Philox stream */
    double boltz;              /* Boltzmann constant in kJ/mol/K; 0 = OpenMM's BOLTZ (8.31446261815324e-3) */
} constv_config;

/* ---- life cycle ------------------------------------------------------------------------------- */

/* Replaces the checks in CudaIntegrateConstVLangevinStepKernel::initialize, Kernels.cpp:76-96:
 * numCells even, N % numCells == 0, every particle with index >= N/numCells massless, and
 * zmax = (cell_size < 0 ? box_z/numCells : cell_size) with zmax*numCells == box_z exactly.
 * masses: N doubles in original particle order (system.getParticleMass).  Writes *zmax_out.
 * Failure: CONSTV_ERR_SYSTEM with the reference's message.  Pure host code (no device needed). */
CONSTV_API int constv_validate_system(int32_t num_atoms, int32_t num_cells, const double* masses,
                                      double box_z, double cell_size, double* zmax_out);

/* Replaces the allocation half of initialize (params + invAtomOrder, Kernels.cpp:71,99) and the
 * RNG seeding (:64).  invAtomOrder starts as the identity (the reference leaves it uninitialised). */
CONSTV_API int constv_create(const constv_config* config, constv_slab** out);

/* Replaces ~CudaIntegrateConstVLangevinStepKernel, Kernels.cpp:53-59. */
CONSTV_API int constv_destroy(constv_slab* slab);

/* Hands the library the device arrays the reference fetches from CudaContext /
 * CudaIntegrationUtilities on every execute (Kernels.cpp:147-166: getVelm, getForce, getPosDelta,
 * getPosq, getPosqCorrection, getAtomIndexArray).  The arrays stay owned by the caller.
 * posq_correction is required for mixed precision; pos_delta only for the split path;
 * atom_index NULL means identity order. */
CONSTV_API int constv_bind_arrays(constv_slab* slab, void* posq, void* posq_correction, void* velm,
                                  void* force, void* pos_delta, const int32_t* atom_index);

/* Alternative to constv_bind_arrays for hosts without device arrays of their own: the library
 * allocates posq/(posq_correction)/velm/force/pos_delta in HBM and binds them. */
CONSTV_API int constv_alloc_arrays(constv_slab* slab);

/* Device addresses of the bound arrays (any output pointer may be NULL). */
CONSTV_API int constv_get_arrays(constv_slab* slab, void** posq, void** posq_correction, void** velm,
                                 void** force, void** pos_delta);

/* ---- per-step entry points -------------------------------------------------------------------- */

/* Replaces the parameter block of execute, Kernels.cpp:113-137: kT = BOLTZ*T,
 * vscale = exp(-dt*friction), fscale = friction == 0 ? dt : (1-vscale)/friction,
 * noisescale = sqrt(kT*(1-vscale^2)), cached on (T, friction, dt), rounded to float for single
 * precision (:98-103).  Also covers setNextStepSize (:112). */
CONSTV_API int constv_set_parameters(constv_slab* slab, double temperature, double friction,
                                     double step_size);

/* out[0..3] = vscale, fscale, noisescale, dt as the kernels see them (after rounding to `mixed`). */
CONSTV_API int constv_get_parameters(const constv_slab* slab, double out[4]);

/* Replaces the reorderInverseAtomOrderIndexes launch, Kernels.cpp:139-142 / Langevin.cu:170-177:
 * invAtomOrder[atom_index[slot]] = slot.  Call when cu.getAtomsWereReordered(). */
CONSTV_API int constv_update_inverse_order(constv_slab* slab, void* stream);

/* Snapshot posq[image slot].w of every image into the side array used with
 * CONSTV_FLAG_CACHE_IMAGE_CHARGE (the value Langevin.cu:153 re-reads every step). */
CONSTV_API int constv_refresh_image_charges(constv_slab* slab, void* stream);

/* Replaces the integrateConstVLangevinPart1 launch, Kernels.cpp:146-149 / Langevin.cu:7-28.
 * random4 != NULL: consume the caller's float4 Gaussian pool at random4[random_index + slot]
 * (exactly the reference's addressing; OpenMM's integration.getRandom()).  random4 == NULL: draw
 * from the built-in counter-based Philox4x32-10 stream keyed (seed; original atom index, step). */
CONSTV_API int constv_step_part1(constv_slab* slab, const void* random4, uint32_t random_index,
                                 void* stream);

/* Replaces the integrateConstVLangevinPart2 + updateImageParticlePositions launches,
 * Kernels.cpp:158-166 / Langevin.cu:34-75,138-164, plus the image zeroing, and advances time and
 * step count (Kernels.cpp:170-171).  The caller applies constraints to pos_delta between
 * constv_step_part1 and this call (integration.applyConstraints, Kernels.cpp:153). */
CONSTV_API int constv_step_part2_mirror(constv_slab* slab, void* stream);

/* Replaces the updateImageParticlePositions launch alone, Kernels.cpp:165-166 (and :346-347 for the Drude
 * integrator) / Langevin.cu:138-164: every charged real atom's images are rewritten from its current posq;
 * identity order or, with an atom_index bound, through invAtomOrder.  No velocity, force or step-counter
 * change.  The reference runs this pass AFTER integration.computeVirtualSites() (Kernels.cpp:161, :342): a
 * host whose System has virtual sites calls the step entry point, computeVirtualSites, then constv_mirror,
 * so that the image of a charged virtual site follows the site's new position (the rewrite is idempotent
 * for every other atom). */
CONSTV_API int constv_mirror(constv_slab* slab, void* stream);

/* The whole of execute (Kernels.cpp:139-172) for a system without constraints in ONE kernel:
 * part 1, part 2, the image rewrite and the image zeroing, each atom record crossing HBM once;
 * pos_delta never leaves registers.  Same random4 / Philox convention as constv_step_part1. */
CONSTV_API int constv_step_fused(constv_slab* slab, const void* random4, uint32_t random_index,
                                 void* stream);

/* One launch that advances n independent slabs (an ensemble of Contexts) by one fused step each;
 * all slabs must share device, precision and identity order.  Philox noise only. */
CONSTV_API int constv_step_fused_batched(constv_slab* const* slabs, int32_t n, void* stream);

/* Replaces computeKineticEnergy, Kernels.cpp:175-177 (OpenMM's
 * integration.computeKineticEnergy(0.5*dt)): KE = 0.5*sum m|v + time_shift*F/m|^2 over massive
 * atoms, accumulated in double.  On a slab with configured rigid molecules (constv_settle_configure) the
 * shifted velocities of the molecule atoms are first projected onto the constraints, as OpenMM's
 * computeKineticEnergy does for a constrained System.  Synchronises `stream` and writes *ke_out on the host. */
CONSTV_API int constv_kinetic_energy(constv_slab* slab, double time_shift, double* ke_out,
                                     void* stream);

/* Same, asynchronous: the double lands in device memory at ke_device (e.g. the buffer a scalar
 * NCCL all-reduce then sums across the range partition). */
CONSTV_API int constv_kinetic_energy_device(constv_slab* slab, double time_shift, double* ke_device,
                                            void* stream);

/* Writes the float4 Gaussians the Philox path would consume at `step` for slots [0, N) to the
 * device array out4 (the stream that stands in for integration.prepareRandomNumbers/getRandom,
 * Kernels.cpp:146,148).  For parity tests and for checkpoint verification. */
CONSTV_API int constv_generate_noise(constv_slab* slab, uint64_t step, void* out4, void* stream);

/* Debug / parity entry: the raw Philox4x32-10 block (four uint32 words) behind those Gaussians, for slots
 * [0, N): counter = (atom lo, atom hi, step lo, step hi) with atom = global_atom_offset + original atom
 * index, key = (seed lo, seed hi).  words4 = device uint32[4*N].  Integer work: bit-exact against the
 * Random123 known-answer vectors and the oracle (tests/test_gpu_philox.py). */
CONSTV_API int constv_generate_philox_words(constv_slab* slab, uint64_t step, void* words4, void* stream);

/* ---- bookkeeping (cu.getTime/setTime, getStepCount/setStepCount, Kernels.cpp:170-171) -------- */

/* The step counter lives in device memory so that CUDA-graph replays advance the Philox stream;
 * reading it synchronises the device.  It is the only persistent state to checkpoint besides the
 * seed. */
CONSTV_API int constv_get_step_count(constv_slab* slab, uint64_t* count);
CONSTV_API int constv_set_step_count(constv_slab* slab, uint64_t count);
CONSTV_API int constv_get_time(const constv_slab* slab, double* time);
CONSTV_API int constv_set_time(constv_slab* slab, double time);

/* ---- host-buffer convenience (the reference-facing call with HOST memory) -------------------- */

/* Copy whole arrays between host buffers (layouts above, any may be NULL to skip) and the bound
 * device arrays.  Synchronous with respect to `stream`. */
CONSTV_API int constv_upload_state(constv_slab* slab, const void* posq, const void* posq_correction,
                                   const void* velm, const int64_t* force, void* stream);
CONSTV_API int constv_download_state(constv_slab* slab, void* posq, void* posq_correction,
                                     void* velm, int64_t* force, void* stream);

/* One step driven from host memory (what `integrator.step(1)` followed by a state query costs a
 * host that keeps its buffers on the CPU): uploads this step's forces for the real atoms
 * (host_force_real: 3 planes of num_real int64, plane stride num_real; NULL = keep the device
 * buffer), runs the fused step, and downloads whichever results are asked for: the kinetic energy
 * at the half-step shift 0.5*dt (host_ke_out, the scalar getState(getEnergy=True) reports), the
 * updated posq (host_posq_out, all N atoms) and velm (host_velm_out).  Pinned host memory makes the
 * copies asynchronous; the call returns after everything requested has landed. */
CONSTV_API int constv_step_host(constv_slab* slab, const int64_t* host_force_real, void* host_posq_out,
                                void* host_velm_out, double* host_ke_out, void* stream);

/* Pipelined form of constv_step_host for hosts that produce forces on the CPU every step: submit
 * step k+1 before collecting step k, and the PCIe upload of step k+1's forces (on an internal copy
 * stream, into a double-buffered staging area) overlaps step k's kernels and read-back.  At most two
 * tickets may be outstanding.  host_force_real: 3 planes of num_real values (plane stride num_real),
 * pinned memory, either OpenMM's int64 fixed point (CONSTV_FORCE_FIXED64) or float32 in kJ/mol/nm
 * (CONSTV_FORCE_FLOAT32, half the PCIe bytes), which the device converts exactly as OpenMM's force
 * kernels fill the buffer: (long long)(f * 2^32), truncating.  The buffer must stay valid until the
 * ticket is collected.  collect() blocks until that step's kinetic energy (half-step shift, as
 * constv_step_host) has landed and returns it. */
enum { CONSTV_FORCE_FIXED64 = 0, CONSTV_FORCE_FLOAT32 = 1 };
CONSTV_API int constv_step_host_submit(constv_slab* slab, const void* host_force_real, int32_t force_format,
                                       uint64_t* ticket_out, void* stream);
CONSTV_API int constv_step_host_collect(constv_slab* slab, uint64_t ticket, double* host_ke_out);

/* constv_step_host_submit that also reads the step's state back (what a trajectory frame per step needs; the
 * reference downloads positions only when a reporter asks, through OpenMM's getState: DCDReporter and
 * context.getState(getPositions=True), example/nacl_1m_spce/nacl_constv.py:156,172,198): after the step, posq of
 * all N atoms (host_posq_out) and / or velm (host_velm_out) are copied to pinned host memory on a second internal
 * stream, i.e. in the other PCIe direction, while the NEXT ticket's forces upload; the next step's kernels wait
 * for the read-back (it copies straight out of the bound arrays).  Either pointer may be NULL; with both NULL this
 * is constv_step_host_submit.  The host buffers belong to the ticket until it is collected: alternate two.
 * collect() returns once the kinetic energy and the requested arrays have landed.  Collect every ticket before
 * calling any other stepping entry point on the slab. */
CONSTV_API int constv_step_host_submit_state(constv_slab* slab, const void* host_force_real, int32_t force_format,
                                             void* host_posq_out, void* host_velm_out, uint64_t* ticket_out,
                                             void* stream);

/* ---- chains: overlapping consecutive steps of one slab ---------------------------------------- */

/* A step of the reference is a sequence of whole-slab launches on one stream (Kernels.cpp:146-166), so
 * step k+1 cannot touch memory before the last block of step k has retired: every step pays its own
 * ramp-up and drain (~2 us of a 12 us step at 1 M atoms).  The update of an atom depends on nothing but
 * its own record, so a slab may be cut into `num_chains` contiguous ranges of real atoms (boundaries on
 * 256-atom tiles), each with its own device step counter; constv_step_fused_chain advances ONE range and
 * may be issued on a different stream per chain, so that the drain of one chain's step overlaps the ramp
 * of another's.  Results are bit-identical to the single launch (the Philox key is the atom index and
 * the step index, not the launch).  constv_step_fused on a chained slab launches every chain in turn on
 * the given stream.  Identity slot order and the plain-load kernel only.  num_chains in 1..16 (clamped
 * to the number of tiles); setting it synchronises the device.  The rigid-molecule and Drude steps have chain
 * entry points of their own (constv_step_fused_settle_chain, constv_drude_step_fused_chain). */
CONSTV_API int constv_set_chains(constv_slab* slab, int32_t num_chains);
CONSTV_API int constv_get_chains(const constv_slab* slab, int32_t* num_chains);
CONSTV_API int constv_step_fused_chain(constv_slab* slab, int32_t chain, const void* random4,
                                       uint32_t random_index, void* stream);

/* num_steps fused steps (Philox noise) of one slab in one call: what `integrator.step(n)` is for a host whose
 * forces do not change between the steps (replayed or constant forces, benchmarks).  On a chained slab the
 * library forks onto one internal stream per chain, issues each chain's num_steps launches in order and joins
 * back into `stream`, so the overlap of constv_set_chains is had without managing streams; the call can be
 * captured in a CUDA graph.  Asynchronous. */
CONSTV_API int constv_step_fused_multi(constv_slab* slab, int32_t num_steps, void* stream);

/* ---- ConstVDrudeLangevinIntegrator ------------------------------------------- */
/* "Drude.cu" = platforms/cuda/src/kernels/constVDrudeLangevin.cu.  The same slab object carries the
 * Drude integrator: create / bind as above, then configure the pair list and use the constv_drude_*
 * parameter and step calls instead of constv_set_parameters / constv_step_*. */

/* Replaces the pair identification of CudaIntegrateConstVDrudeLangevinStepKernel::initialize,
 * Kernels.cpp:213-230: pair_particles = num_pairs x (Drude particle, parent) as
 * DrudeForce::getParticleParameters returns (p, p1), HOST memory; every other real atom is an
 * "ordinary" particle.  The reference uses these indices directly as slots (it never translates them
 * through the atom order; OpenMM only swaps identical molecules), and so does this library.
 * Both particles must be distinct real atoms and belong to one pair only. */
CONSTV_API int constv_drude_configure(constv_slab* slab, int32_t num_pairs, const int32_t* pair_particles);

/* numNormal + 2*numPairs: the count the reference hands to prepareRandomNumbers (Kernels.cpp:307).
 * Pool addressing is the reference's: ordinary particle `index` reads random4[random_index + index]
 * (Drude.cu:18), pair i reads random4[random_index + numNormal + 2i] and the next entry (Drude.cu:30,46-47). */
CONSTV_API int constv_drude_random_count(const constv_slab* slab, int32_t* count);

/* Replaces the coefficient block of execute, Kernels.cpp:261-301: the (temperature, friction) thermostat
 * of ordinary particles and of each pair's centre of mass, the (drude_temperature, drude_friction)
 * thermostat of each pair's relative motion, max_drude_distance (hard wall, <= 0 disables it) and
 * hardwallscaleDrude = sqrt(kB*drude_temperature); rounded to float in single precision. */
CONSTV_API int constv_drude_set_parameters(constv_slab* slab, double temperature, double friction,
                                           double drude_temperature, double drude_friction,
                                           double max_drude_distance, double step_size);

/* out[0..8] = vscale, fscale/2^32, noisescale, vscaleDrude, fscaleDrude/2^32, noisescaleDrude,
 * maxDrudeDistance, hardwallscaleDrude, dt as the kernels see them. */
CONSTV_API int constv_drude_get_parameters(const constv_slab* slab, double out[9]);

/* Replaces the integrateConstVDrudeLangevinPart1 launch, Kernels.cpp:307-311 / Drude.cu:5-66.
 * random4 == NULL: Philox; an ordinary particle draws its own block, a pair the Drude particle's block
 * for the centre of mass and the parent's for the relative motion. */
CONSTV_API int constv_drude_step_part1(constv_slab* slab, const void* random4, uint32_t random_index,
                                       void* stream);

/* Replaces integrateConstVDrudeLangevinPart2 + applyHardWallConstraints + updateImageParticlePositions
 * (Kernels.cpp:321-344 / Drude.cu:72-211, Langevin.cu:138-164) in ONE kernel, plus the image zeroing;
 * advances time and step count.  The caller applies constraints to pos_delta before it (Kernels.cpp:315). */
CONSTV_API int constv_drude_step_part2(constv_slab* slab, void* stream);

/* The whole of execute (Kernels.cpp:303-349) for a system without constraints in ONE kernel. */
CONSTV_API int constv_drude_step_fused(constv_slab* slab, const void* random4, uint32_t random_index,
                                       void* stream);

/* One chain of that step (constv_set_chains; identity slot order, the staged kernel), as
 * constv_step_fused_settle_chain. */
CONSTV_API int constv_drude_step_fused_chain(constv_slab* slab, int32_t chain, const void* random4,
                                             uint32_t random_index, void* stream);

/* applyHardWallConstraints alone (Kernels.cpp:336-340 / Drude.cu:108-211), thread <-> pair, for hosts
 * that run their own part 2.  No-op when max_drude_distance <= 0. */
CONSTV_API int constv_drude_hardwall(constv_slab* slab, void* stream);

/* ---- rigid three-site molecules inside the step ------------------------------ */
/* The reference hands posDelta to OpenMM's constraint solver between part 1 and part 2
 * (integration.applyConstraints, Kernels.cpp:153).  For its example system -- SPC/E water with
 * constraints=AllBonds, example/nacl_1m_spce/nacl_constv.py:56 -- every constraint belongs to a rigid 3-site
 * molecule, which OpenMM's CUDA platform solves analytically (SETTLE, Miyamoto & Kollman 1992).  These
 * entry points take that case into the step: part 1, SETTLE, part 2, the image rewrite and the zeroing in ONE
 * kernel, posDelta never leaving registers.  The solver is OpenMM's, not the reference's: it is restated from
 * the published algorithm (parity unpinned). */

/* The rule by which OpenMM's CudaIntegrationUtilities picks the constraints SETTLE solves [recalled]: atoms
 * with exactly two constraints each, closing a triangle; two of the three distances equal, the atom between
 * them is the apex.  Pure host code.  cluster_atoms (3 per molecule: apex, b, c) and cluster_params (2 per
 * molecule: d(apex-b) = d(apex-c), d(b-c)) may be NULL to count only; they need room for num_constraints/3
 * molecules.  *num_other_constraints_out = constraints left for a general solver (0 = the fused step applies).
 * Three unequal distances: CONSTV_ERR_SYSTEM "Two of the three distances constrained with SETTLE must be the same." */
CONSTV_API int constv_settle_find_clusters(int32_t num_atoms, int32_t num_constraints, const int32_t* atom1,
                                           const int32_t* atom2, const double* distance, int32_t* cluster_atoms,
                                           float* cluster_params, int32_t* num_clusters_out,
                                           int32_t* num_other_constraints_out);

/* Upload the molecule list (HOST arrays, layout as above; atoms are slots of real atoms, used directly as
 * OpenMM's solver uses them; OpenMM only swaps identical molecules, so a slot keeps its role). */
CONSTV_API int constv_settle_configure(constv_slab* slab, int32_t num_clusters, const int32_t* cluster_atoms,
                                       const float* cluster_params);

/* The whole of execute (Kernels.cpp:139-172) INCLUDING the constraint hand-off at :153, for a system whose
 * constraints are all in the configured molecules, in one kernel.  random4 / Philox as constv_step_part1. */
CONSTV_API int constv_step_fused_settle(constv_slab* slab, const void* random4, uint32_t random_index,
                                        void* stream);

/* One chain of that step (constv_set_chains; identity slot order, the staged kernel): the chain's share of the
 * step's tiles with its own step counter, to be issued on a stream of its own.  constv_step_fused_settle on a
 * chained slab issues every chain in turn on the given stream. */
CONSTV_API int constv_step_fused_settle_chain(constv_slab* slab, int32_t chain, const void* random4,
                                              uint32_t random_index, void* stream);

/* One launch (per group of 8) that advances n independent rigid-molecule slabs -- an ensemble of Contexts -- by
 * one step each; all slabs must share device and precision, be in identity slot order, unchained, and take the
 * staged kernel.  Philox noise only. */
CONSTV_API int constv_step_fused_settle_batched(constv_slab* const* slabs, int32_t n, void* stream);

/* The solver alone on pos_delta, thread <-> molecule: what applyConstraints does for these molecules, for
 * the split path constv_step_part1 | constv_settle_positions | constv_step_part2_mirror. */
CONSTV_API int constv_settle_positions(constv_slab* slab, void* stream);

/* ---- diagnostics ------------------------------------------------------------------------------ */

CONSTV_API const char* constv_last_error(void);
CONSTV_API int constv_version(void);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
CONSTV_API uint64_t constv_launch_count(void);
/* Name of the kernel behind the calling thread's most recent launch, where the library had a choice ("" otherwise):
 * the rigid-molecule step's constv_settle_step_{tiled,bulk,ws,tiled_batched,bulk_batched}_kernel, constv_fused_step_tma_kernel.
 * Diagnostics and tests (which variant did constv_set_kernel_variant / the automatic rule pick). */
CONSTV_API const char* constv_last_kernel(void);
/* Tuning knobs of the fused kernel: real/image pairs per thread (1, 2 or 4) and the grid:
 * ctas_per_sm = -1 launches a one-shot grid (one tile per CTA, the default), 1..32 a persistent grid
 * of that many CTAs per SM; 0 keeps the current value. */
CONSTV_API int constv_set_launch_shape(constv_slab* slab, int32_t pairs_per_thread, int32_t ctas_per_sm);

/* Which fused kernel constv_step_fused launches: 0 = automatic (currently the plain-load one-shot
 * kernel, which measured faster than the TMA ring on B200), 1 = the plain-load kernel, 2 = the
 * TMA-pipelined kernel (error unless the slab is eligible: identity order, Philox noise,
 * CONSTV_FLAG_CACHE_IMAGE_CHARGE, even P).  For the rigid-molecule step (constv_step_fused_settle*): 0 = automatic
 * (the staged kernel; its stage is filled by bulk copies issued by one thread for a chain's launch and for launches of
 * at least two waves of CTAs, by per-thread asynchronous copies otherwise), 1 = the gather kernel, 3 = the staged
 * kernel with its stage filled through registers (round 1's), 4 = the warp-specialised TMA pipeline, 6 = always the
 * bulk stage-in, 7 = always the per-thread asynchronous copies (4 and 6 need identity order, Philox noise,
 * CONSTV_FLAG_CACHE_IMAGE_CHARGE, num_cells == 2 and P a multiple of 4, else the cp.async-staged kernel runs); same
 * results bit for bit.  For the Drude step (constv_drude_step_fused*): 0 = automatic (the
 * staged kernel whose warps draw chunks of work items), 1 = the thread-per-slot gather kernel, 5 = the staged kernel with
 * one work item per thread (round 1's); same results bit for bit.
 * (block_size, stages) in {(128,6), (256,2), (256,3), (256,4), (512,2), (512,3)} shape the TMA ring, ctas_per_sm caps the persistent
 * grid (0 = as many as shared memory allows); 0 keeps the current block_size / stages. */
CONSTV_API int constv_set_kernel_variant(constv_slab* slab, int32_t variant, int32_t block_size,
                                         int32_t stages, int32_t ctas_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* CONSTV_B200_H_ */
