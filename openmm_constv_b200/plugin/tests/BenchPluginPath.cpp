// Times the drop-in path: one CudaIntegrateConstVLangevinStepKernel::execute per step, created through the
// platform's kernel factory exactly as ConstVLangevinIntegrator::initialize does, with a cache-sweeping
// device operation between two steps standing in for the force kernels a host (OpenMM) runs there.  Unlike
// bench.py's headline there is no CUDA graph, no second stream and nothing of the next step to overlap with:
// this is what a Context stepping through the plugin gets.
//
//   BenchPluginPath [--atoms N] [--steps S] [--precision single|mixed|double] [--reordered] [--sweep-mb M]
//
// Reported (one JSON line), for three sweeps between two steps --
//   forces  a kernel that copies into the force buffer (24 B per atom): what the cheapest possible force
//           kernel does; posq / velm may then still be L2 hits
//   copy    a kernel that copies 128 MB onto another 128 MB: reads and writes more than the 126 MB L2
//           holds, like the force evaluation of a large system; nothing of the slab survives in L2
//   read    a kernel that only READS 256 MB: the L2 ends up full of clean lines of something else
// -- two numbers per step:
//   marginal   (T[sweep + execute loop] - T[sweep-only loop]) / steps, medians of 5 passes: launch gap, ramp,
//              transfer and drain of the step kernel PLUS what the step costs its neighbour (the write-back of
//              the step's 37 MB of dirty lines happens while the next sweep runs and slows it down)
//   bracketed  CUDA events around each execute(): the step's own duration and launch gap
// --reordered applies one slot permutation of the kind cu.reorderAtoms() produces (whole molecules trade
// places inside windows of 1024 molecules) so that execute() addresses images through atomIndex /
// invAtomOrder (constVLangevin.cu:142-158).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "CudaContext.h"
#include "CudaPlatform.h"
#include "openmm/ConstVKernels.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/Context.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;
using std::vector;

extern "C" void registerConstVCudaKernelFactories();
extern "C" void benchSweepRead(const void* p, size_t bytes, unsigned int* sink, cudaStream_t stream);
extern "C" void benchSweepCopy(void* dst, const void* src, size_t bytes, cudaStream_t stream);

#define CU(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            std::exit(2);                                                                  \
        }                                                                                  \
    } while (0)

static double median(vector<double> v) {
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main(int argc, char** argv) {
    int atoms = 1000000, steps = 300;
    size_t sweepMB = 256;
    std::string precision = "single";
    bool reordered = false, rigidWater = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--atoms" && i + 1 < argc) atoms = std::atoi(argv[++i]);
        else if (a == "--steps" && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (a == "--precision" && i + 1 < argc) precision = argv[++i];
        else if (a == "--sweep-mb" && i + 1 < argc) sweepMB = (size_t) std::atoi(argv[++i]);
        else if (a == "--reordered") reordered = true;
        else if (a == "--rigid-water") rigidWater = true;  // constraints=AllBonds + rigidWater (nacl_constv.py:56): solved inside the step
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    try {
        registerConstVCudaKernelFactories();
        // the example's composition (nacl_1m_eq.pdb): ~1 % ions, 82 % SPC/E water, 17 % massless neutral wall atoms
        const int R = atoms / 2;
        const int nWall = (int) std::lround(R * (1280.0 / 7494.0));
        const int nIon = (int) std::lround(R * (76.0 / 7494.0));
        const int nWater = (R - nWall - nIon) / 3 * 3;
        const int nIonAll = R - nWall - nWater;
        vector<double> mass(2 * (size_t) R, 0.0), charge(2 * (size_t) R, 0.0);
        for (int i = 0; i < nIonAll; i++) { mass[i] = i % 2 ? 35.45 : 22.98977; charge[i] = i % 2 ? -1.0 : 1.0; }
        for (int i = 0; i < nWater; i++) {
            mass[nIonAll + i] = i % 3 ? 1.008 : 15.9994;
            charge[nIonAll + i] = i % 3 ? 0.4238 : -0.8476;
        }
        for (int i = 0; i < R; i++) charge[R + i] = -charge[i];
        const double zmax = 4.2183, scale = std::sqrt(R / 7494.0), lx = 3.9352 * scale, ly = 4.26 * scale;
        System system;
        for (double m : mass) system.addParticle(m);
        system.setDefaultPeriodicBoxVectors(Vec3(lx, 0, 0), Vec3(0, ly, 0), Vec3(0, 0, 2 * zmax));
        const double dOH = 0.1, theta = 109.47 * M_PI / 180.0, dHH = 2 * dOH * std::sin(0.5 * theta);  // SPC/E
        if (rigidWater)
            for (int k = 0; k < nWater / 3; k++) {
                const int o = nIonAll + 3 * k;
                system.addConstraint(o, o + 1, dOH);
                system.addConstraint(o, o + 2, dOH);
                system.addConstraint(o + 1, o + 2, dHH);
            }
        ConstVLangevinIntegrator integrator(300.0, 1.0, 0.001);
        integrator.setRandomNumberSeed(2024);
        Platform& platform = Platform::getPlatformByName("CUDA");
        std::map<std::string, std::string> properties;
        properties["Precision"] = precision;
        Context context(system, integrator, platform, properties);
        std::mt19937_64 gen(12345);
        std::uniform_real_distribution<double> uni(0.0, 1.0);
        vector<Vec3> pos(2 * (size_t) R);
        for (int i = 0; i < R; i++) {
            const bool wall = i >= nIonAll + nWater;
            const double z = wall ? ((i & 1) ? zmax : 0.0) : (0.05 + 0.9 * uni(gen)) * zmax;
            pos[i] = Vec3(uni(gen) * lx, uni(gen) * ly, z);
            const int w = i - nIonAll;
            if (rigidWater && w >= 0 && w < nWater && w % 3 != 0) {  // the hydrogens on the SPC/E geometry around their oxygen
                const Vec3& o = pos[i - w % 3];
                pos[i] = w % 3 == 1 ? Vec3(o[0] + dOH, o[1], o[2]) : Vec3(o[0] + dOH * std::cos(theta), o[1] + dOH * std::sin(theta), o[2]);
            }
            pos[R + i] = Vec3(pos[i][0], pos[i][1], charge[i] != 0 ? -pos[i][2] : -pos[i][2] + 0.001);
        }
        context.setPositions(pos);
        context.setCharges(charge);
        context.setVelocitiesToTemperature(300.0, 7);

        CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getImpl().getPlatformData())->contexts[0];
        const int P = cu.getPaddedNumAtoms();
        // injected forces in OpenMM's fixed-point layout, images included (so that the zeroing is real work)
        vector<long long> force(3 * (size_t) P, 0);
        std::normal_distribution<double> normal(0.0, 1000.0);
        for (int k = 0; k < 3; k++)
            for (int i = 0; i < 2 * R; i++) force[(size_t) k * P + i] = (long long) std::llround(normal(gen) * 4294967296.0);
        cu.getForce().uploadBytes(force.data());
        void* spareForce = NULL;
        CU(cudaMalloc(&spareForce, 24 * (size_t) P));
        CU(cudaMemcpy(spareForce, force.data(), 24 * (size_t) P, cudaMemcpyHostToDevice));
        void* sweep = NULL;
        CU(cudaMalloc(&sweep, sweepMB << 20));
        CU(cudaMemset(sweep, 0, sweepMB << 20));
        unsigned int* sink = NULL;
        CU(cudaMalloc(&sink, 4));

        // the kernel, made and initialised the way ConstVLangevinIntegrator::initialize does it
        Kernel kernel = platform.createKernel(IntegrateConstVLangevinStepKernel::Name(), context.getImpl());
        IntegrateConstVLangevinStepKernel& step = kernel.getAs<IntegrateConstVLangevinStepKernel>();
        step.initialize(system, integrator);
        step.execute(context.getImpl(), integrator);
        if (reordered) {
            vector<int> order(2 * (size_t) R);
            for (int i = 0; i < 2 * R; i++) order[i] = i;
            const int nMol = nWater / 3, window = 1024;
            std::mt19937 pgen(7);
            for (int lo = 0; lo < nMol; lo += window) {
                const int hi = std::min(nMol, lo + window);
                vector<int> w(hi - lo);
                for (int j = 0; j < hi - lo; j++) w[j] = lo + j;
                std::shuffle(w.begin(), w.end(), pgen);
                for (int j = 0; j < hi - lo; j++)
                    for (int k = 0; k < 3; k++)
                        for (int cell = 0; cell < 2; cell++)
                            order[(size_t) cell * R + nIonAll + 3 * (lo + j) + k] = cell * R + nIonAll + 3 * w[j] + k;
            }
            cu.setPendingReorder(order);
            step.execute(context.getImpl(), integrator);  // cu.reorderAtoms() at its end applies the permutation
            cu.getForce().uploadBytes(force.data());
        }
        for (int i = 0; i < 5; i++) step.execute(context.getImpl(), integrator);
        CU(cudaDeviceSynchronize());

        cudaStream_t stream = 0;
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        // sweeps: 0 forces = D2D copy into the force buffer (24 B per atom); 1 copy = D2D copy of half the sweep buffer
        // onto the other half (reads and writes > L2, like a force evaluation of a large system);
        // 2 read = a kernel that only reads the sweep buffer
        const size_t half = (sweepMB << 20) / 2;
        auto sweepOp = [&](int mode) {
            if (mode == 0) benchSweepCopy((void*) cu.getForce().getDevicePointer(), spareForce, 24 * (size_t) P, stream);
            else if (mode == 1) benchSweepCopy((char*) sweep + half, sweep, half, stream);
            else benchSweepRead(sweep, sweepMB << 20, sink, stream);
        };
        auto pass = [&](int mode, bool withStep) {
            CU(cudaEventRecord(e0, stream));
            for (int i = 0; i < steps; i++) {
                sweepOp(mode);
                if (withStep) step.execute(context.getImpl(), integrator);
            }
            CU(cudaEventRecord(e1, stream));
            CU(cudaEventSynchronize(e1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            return (double) ms * 1e3 / steps;
        };
        // the step alone between two events (its own duration + launch gap; what it costs its neighbours is not in it)
        vector<cudaEvent_t> ev(2 * (size_t) steps);
        for (auto& e : ev) CU(cudaEventCreate(&e));
        auto bracketed = [&](int mode) {
            for (int i = 0; i < steps; i++) {
                sweepOp(mode);
                CU(cudaEventRecord(ev[2 * i], stream));
                step.execute(context.getImpl(), integrator);
                CU(cudaEventRecord(ev[2 * i + 1], stream));
            }
            CU(cudaEventSynchronize(ev[2 * (size_t) steps - 1]));
            vector<double> t;
            for (int i = 0; i < steps; i++) {
                float ms = 0;
                CU(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
                t.push_back(ms * 1e3);
            }
            std::sort(t.begin(), t.end());
            double sum = 0;
            int n = 0;
            for (int i = steps / 10; i < steps - steps / 10; i++, n++) sum += t[i];  // mean of the middle 80 %
            return sum / n;
        };
        double us[3], sweepUs[3], alone[3];
        for (int mode = 0; mode < 3; mode++) {
            vector<double> with, without;
            pass(mode, true);
            for (int r = 0; r < 5; r++) {
                without.push_back(pass(mode, false));
                with.push_back(pass(mode, true));
            }
            sweepUs[mode] = median(without);
            us[mode] = median(with) - sweepUs[mode];
            alone[mode] = bracketed(mode);
        }

        // algorithmic bytes of one step (DESIGN.md "bytes per pair"), numCells = 2
        const double rb = precision == "double" ? 8 : 4, mb = precision == "single" ? 4 : 8;
        const double posB = 4 * rb * (precision == "mixed" ? 2 : 1), velB = 4 * mb;
        const long long massive = nIonAll + nWater, charged = nIonAll + nWater;
        double bytes = (double) R * (velB + posB) + (double) massive * (24 + velB + posB) + (double) charged * (rb + posB) +
                       (double) R * (velB + 24);
        if (reordered) bytes += 8.0 * R + 4.0 * charged;  // atomIndex of both slots, invAtomOrder of the image
        const char* names[3] = {"forces", "copy", "read"};
        const char* what[3] = {"a kernel copying into the force buffer (24 B/atom) between steps", "a kernel copying 128 MB onto 128 MB between steps",
                               "a kernel reading 256 MB between steps"};
        std::printf("{\"path\": \"CudaIntegrateConstVLangevinStepKernel::execute via the kernel factory (OpenMM shim)\", "
                    "\"atoms\": %d, \"precision\": \"%s\", \"reordered_slots\": %s, \"rigid_water\": %s, \"steps\": %d, \"launches_per_step\": 1, "
                    "\"algorithmic_bytes_per_step\": %.0f, ",
                    atoms, precision.c_str(), reordered ? "true" : "false", rigidWater ? "true" : "false", steps, bytes);
        for (int mode = 0; mode < 3; mode++)
            std::printf("\"%s\": {\"marginal_us\": %.3f, \"marginal_frac_8TBs\": %.4f, \"bracketed_us\": %.3f, \"bracketed_frac_8TBs\": %.4f, "
                        "\"sweep\": \"%s, %.1f us\"}, ",
                        names[mode], us[mode], bytes / us[mode] * 1e-3 / 8000.0, alone[mode], bytes / alone[mode] * 1e-3 / 8000.0, what[mode],
                        sweepUs[mode]);
        if (rigidWater) {  // sanity of what was timed: the constraints hold after all those steps (original order, whatever the slots)
            State state = context.getState(State::Positions);
            const vector<Vec3>& x = state.getPositions();
            double worst = 0;
            for (int k = 0; k < nWater / 3; k++) {
                const int o = nIonAll + 3 * k;
                const Vec3 a = x[o + 1] - x[o], b = x[o + 2] - x[o], c = x[o + 2] - x[o + 1];
                worst = std::max(worst, std::fabs(std::sqrt(a.dot(a)) - dOH));
                worst = std::max(worst, std::fabs(std::sqrt(b.dot(b)) - dOH));
                worst = std::max(worst, std::fabs(std::sqrt(c.dot(c)) - dHH));
            }
            // the injected forces are constant and large (sigma 1000 kJ/mol/nm for thousands of steps), so the molecules drift far:
            // the error to expect is a few units in the last place of the largest coordinate
            double far = 0;
            for (int i = 0; i < R; i++)
                for (int k = 0; k < 3; k++) far = std::max(far, std::fabs(x[i][k]));
            std::printf("\"max_constraint_error_nm\": %.3g, \"max_abs_coordinate_nm\": %.3g, ", worst, far);
        }
        std::printf("\"statistic\": \"marginal = median of 5 (sweep+execute) passes minus median of 5 sweep-only passes, per step; "
                    "bracketed = CUDA events around each execute(), mean of the middle 80 %% of the steps\"}\n");
        for (auto& e : ev) cudaEventDestroy(e);
        kernel = Kernel();
        cudaFree(spareForce);
        cudaFree(sweep);
        cudaFree(sink);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
