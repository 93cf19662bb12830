#pragma once

/// \file GpuOutput.h
/// \brief BinaryOutput (.ssf) written straight from the device-resident state of a GpuSolver.
///
/// The reference dumps a snapshot with BinaryOutput::dump(storage, stats) (core/io/Output.cpp:481-582): header, quantity
/// table, then per material its settings and the value / derivative buffers of every quantity. For a run stepped by
/// GpuPredictorCorrector the particle data live on the device; GpuBinaryOutput writes the same file -- same bytes, it can
/// be loaded by the reference's BinaryInput and resumed -- but takes the particle data of every quantity that is mirrored
/// on the device directly from the device planes, through two page-locked staging buffers: the device -> host copy of the
/// next buffer is in flight while the previous one is written to the file, and the host Storage is not touched (no
/// syncToHost, no AoS repack on the host). Everything that is not particle data (header, material settings, ranges,
/// attractors) and quantities the device does not hold come from the Storage.
///
///   reference class replaced          file:line in the reference
///   BinaryOutput (IOutput)            core/io/Output.h:249-345 (format specification), core/io/Output.cpp:481-582

#include "GpuSolver.h"
#include "io/Output.h"

NAMESPACE_SPH_BEGIN

class GpuBinaryOutput : public IOutput {
private:
    GpuSolver& gpu;
    RunTypeEnum runTypeId;
    void* staging[2] = { nullptr, nullptr }; ///< page-locked, sphgpu_host_alloc
    std::size_t stagingBytes = 0;

public:
    GpuBinaryOutput(const OutputFile& fileMask, GpuSolver& gpu, const RunTypeEnum runTypeId = RunTypeEnum::SPH);

    ~GpuBinaryOutput() override;

    /// \param storage The Storage the GpuSolver mirrors (its particle count, materials and quantity set define the file).
    virtual Expected<Path> dump(const Storage& storage, const Statistics& stats) override;
};

NAMESPACE_SPH_END
