#include "GpuOutput.h"
#include "../../include/sphgpu.h"
#include "io/FileSystem.h"
#include "io/Serializer.h"
#include "objects/utility/Streams.h"
#include "quantities/IMaterial.h"
#include "quantities/Attractor.h"
#include "quantities/Quantity.h"
#include "quantities/Storage.h"
#include "system/Statistics.h"
#include <cstring>
#include <vector>

NAMESPACE_SPH_BEGIN

namespace {

/// Device mirror of a Storage quantity: C-ABI quantity id, doubles per particle in the file (0: a Size quantity).
struct Mirror {
    QuantityId id;
    int q;
    int doubles;
    int maxOrder;
};

const Mirror MIRRORS[] = {
    { QuantityId::POSITION, SPHGPU_Q_POSITION, 4, 2 },
    { QuantityId::MASS, SPHGPU_Q_MASS, 1, 0 },
    { QuantityId::PRESSURE, SPHGPU_Q_PRESSURE, 1, 0 },
    { QuantityId::SOUND_SPEED, SPHGPU_Q_SOUND_SPEED, 1, 0 },
    { QuantityId::DENSITY, SPHGPU_Q_DENSITY, 1, 1 },
    { QuantityId::ENERGY, SPHGPU_Q_ENERGY, 1, 1 },
    { QuantityId::DAMAGE, SPHGPU_Q_DAMAGE, 1, 1 },
    { QuantityId::DEVIATORIC_STRESS, SPHGPU_Q_DEVIATORIC_STRESS, 5, 1 },
    { QuantityId::STRESS_REDUCING, SPHGPU_Q_STRESS_REDUCING, 1, 0 },
    { QuantityId::VELOCITY_DIVERGENCE, SPHGPU_Q_VELOCITY_DIVERGENCE, 1, 0 },
    { QuantityId::VELOCITY_GRADIENT, SPHGPU_Q_VELOCITY_GRADIENT, 6, 0 },
    { QuantityId::VELOCITY_ROTATION, SPHGPU_Q_VELOCITY_ROTATION, 4, 0 },
    { QuantityId::STRAIN_RATE_CORRECTION_TENSOR, SPHGPU_Q_CORRECTION_TENSOR, 6, 0 },
    { QuantityId::EPS_MIN, SPHGPU_Q_EPS_MIN, 1, 0 },
    { QuantityId::M_ZERO, SPHGPU_Q_M_ZERO, 1, 0 },
    { QuantityId::EXPLICIT_GROWTH, SPHGPU_Q_EXPLICIT_GROWTH, 1, 0 },
    { QuantityId::N_FLAWS, SPHGPU_Q_N_FLAWS, 0, 0 },
    { QuantityId::FLAG, SPHGPU_Q_FLAG, 0, 0 },
    { QuantityId::NEIGHBOR_CNT, SPHGPU_Q_NEIGHBOR_CNT, 0, 0 },
};

const Mirror* findMirror(const QuantityId id, const Quantity& q) {
    for (const Mirror& m : MIRRORS) {
        if (m.id == id && int(q.getOrderEnum()) <= m.maxOrder) {
            return &m;
        }
    }
    return nullptr;
}

void checkGpu(const int rc) {
    if (rc != SPHGPU_OK) {
        throw IoError("GpuBinaryOutput: " + String::fromAscii(sphgpu_last_error()));
    }
}

/// Fixed 16-character field of the header (Output.h:266-267).
void writeFixedString(const String& s, Serializer<true>& serializer) {
    char field[16];
    for (Size i = 0; i < 16; ++i) {
        field[i] = i < s.size() ? char(s[i]) : '\0';
    }
    serializer.write(field);
}

template <typename TSettingsEntries>
void writeSettings(Serializer<true>& serializer, const TSettingsEntries& settings) {
    serializer.serialize(settings.size());
    for (auto param : settings) {
        serializer.serialize(param.id);
        serializer.serialize(param.value.getTypeIdx());
        forValue(param.value, [&serializer](const auto& value) { serializer.write(value); });
    }
}

/// Buffers of a quantity the device does not hold, from the Storage (value, then the derivatives the order has).
struct HostBuffersVisitor {
    template <typename TValue>
    void visit(const Quantity& q, Serializer<true>& serializer, const IndexSequence& sequence) {
        StaticArray<const Array<TValue>&, 3> buffers = q.template getAll<TValue>();
        const int orders = int(q.getOrderEnum()) + 1;
        for (int o = 0; o < orders; ++o) {
            for (Size i : sequence) {
                serializer.write(buffers[o][i]);
            }
        }
    }
};

} // namespace

GpuBinaryOutput::GpuBinaryOutput(const OutputFile& fileMask, GpuSolver& gpu, const RunTypeEnum runTypeId)
    : IOutput(fileMask)
    , gpu(gpu)
    , runTypeId(runTypeId) {}

GpuBinaryOutput::~GpuBinaryOutput() {
    sphgpu_host_free(staging[0]);
    sphgpu_host_free(staging[1]);
}

Expected<Path> GpuBinaryOutput::dump(const Storage& storage, const Statistics& stats) {
    const Path fileName = paths.getNextPath(stats);
    Outcome dirResult = FileSystem::createDirectory(fileName.parentPath());
    if (!dirResult) {
        return makeUnexpected<Path>("Cannot create directory {}: {}", fileName.parentPath().string(), dirResult.error());
    }
    sphgpu_ctx* ctx = gpu.context(storage);
    const Size particleCnt = storage.getParticleCnt();
    // largest sub-block: a Vector buffer of all particles (Sizes are widened to int64 on the host: 8 bytes as well)
    const std::size_t need = std::max<std::size_t>(std::size_t(particleCnt) * 6 * sizeof(double), 64);
    if (need > stagingBytes) {
        sphgpu_host_free(staging[0]);
        sphgpu_host_free(staging[1]);
        checkGpu(sphgpu_host_alloc(&staging[0], need));
        checkGpu(sphgpu_host_alloc(&staging[1], need));
        stagingBytes = need;
    }

    AutoPtr<FileBinaryOutputStream> file = makeAuto<FileBinaryOutputStream>(fileName);
    FileBinaryOutputStream* raw = &*file; // particle data go to the stream directly, everything else through the serializer
    Serializer<true> serializer(std::move(file));

    // ---- header: 256 bytes (Output.h:258-270) ----
    const Size materialCnt = storage.getMaterialCnt();
    const Size quantityCnt = storage.getQuantityCnt() - int(storage.has(QuantityId::MATERIAL_ID));
    serializer.serialize("SPH", stats.getOr<Float>(StatisticsId::RUN_TIME, 0._f), particleCnt, quantityCnt, materialCnt,
        stats.getOr<Float>(StatisticsId::TIMESTEP_VALUE, 0.1_f), BinaryIoVersion::LATEST);
    writeFixedString(EnumMap::toString(runTypeId), serializer);
    writeFixedString(__DATE__, serializer);
    serializer.serialize(Size(stats.getOr<int>(StatisticsId::WALLCLOCK_TIME, 0)));
    serializer.serialize(storage.getAttractorCnt());
    serializer.addPadding(156);

    // ---- quantity table ----
    Array<QuantityId> ids;
    for (auto i : storage.getQuantities()) {
        if (i.id != QuantityId::MATERIAL_ID) {
            ids.push(i.id);
            serializer.serialize(Size(i.id), Size(i.quantity.getOrderEnum()), Size(i.quantity.getValueEnum()));
        }
    }

    // One sub-block (one buffer of one quantity over the particles of one material) is copied from the device while the
    // previous one is being written: `pending` describes the copy in flight.
    struct Pending {
        int slot = -1;
        std::size_t bytes = 0;
        bool widen = false;
        Size count = 0;
    } pending;
    int nextSlot = 0;
    std::vector<int64_t> widened;
    auto flush = [&]() {
        if (pending.slot < 0) {
            return;
        }
        checkGpu(sphgpu_transfer_sync(ctx));
        if (pending.widen) { // Size quantities: the device holds u32, the file int64
            widened.resize(pending.count);
            const uint32_t* src = static_cast<const uint32_t*>(staging[pending.slot]);
            for (Size k = 0; k < pending.count; ++k) {
                widened[k] = int64_t(src[k]);
            }
            raw->write(ArrayView<const char>(reinterpret_cast<const char*>(widened.data()), pending.count * sizeof(int64_t)));
        } else {
            raw->write(ArrayView<const char>(static_cast<const char*>(staging[pending.slot]), pending.bytes));
        }
        pending.slot = -1;
    };

    const bool hasMaterials = materialCnt > 0;
    for (Size matIdx = 0; matIdx < max(materialCnt, Size(1)); ++matIdx) {
        IndexSequence sequence(0, particleCnt);
        if (hasMaterials) {
            serializer.serialize("MAT", matIdx);
            MaterialView material = storage.getMaterial(matIdx);
            writeSettings(serializer, material->getParams());
            for (QuantityId id : ids) {
                const Interval range = material->range(id);
                serializer.serialize(id, range.lower(), range.upper(), material->minimal(id));
            }
            sequence = material.sequence();
        } else {
            serializer.serialize("NOMAT");
        }
        const Size first = *sequence.begin(), count = *sequence.end() - *sequence.begin();
        serializer.serialize(first, first + count);

        for (auto i : storage.getQuantities()) {
            if (i.id == QuantityId::MATERIAL_ID) {
                continue;
            }
            const Quantity& q = i.quantity;
            const Mirror* m = findMirror(i.id, q);
            if (!m || count == 0) {
                flush(); // keep the file order
                if (count > 0) {
                    HostBuffersVisitor visitor;
                    dispatch(q.getValueEnum(), visitor, q, serializer, sequence);
                }
                continue;
            }
            const int orders = int(q.getOrderEnum()) + 1;
            for (int o = 0; o < orders; ++o) {
                const int slot = nextSlot;
                nextSlot ^= 1;
                // queue the copy of this sub-block, then write the previous one while it is in flight
                checkGpu(sphgpu_download_async(ctx, m->q, o, SPHGPU_LAYOUT_PACKED, staging[slot], first, count));
                checkGpu(sphgpu_download_batch_end(ctx));
                Pending mine;
                mine.slot = slot;
                mine.widen = m->doubles == 0;
                mine.count = count;
                mine.bytes = std::size_t(count) * (m->doubles == 0 ? sizeof(uint32_t) : m->doubles * sizeof(double));
                flush();
                pending = mine;
            }
        }
        flush();
    }
    for (const Attractor& a : storage.getAttractors()) {
        serializer.write(a.position);
        serializer.write(a.velocity);
        serializer.write(a.radius);
        serializer.write(a.mass);
        writeSettings(serializer, a.settings);
    }
    return fileName;
}

NAMESPACE_SPH_END
