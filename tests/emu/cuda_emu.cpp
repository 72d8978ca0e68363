// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the launch loop of the CPU emulation.
#include "cuda_emu.h"

namespace emu {

thread_local idx3 t_thread_idx, t_block_idx;
thread_local Cta *t_cta = nullptr;
thread_local unsigned t_static_off = 0;
idx3 g_block_dim, g_grid_dim;

void launch(idx3 grid, unsigned block_threads, size_t dyn_smem_bytes, unsigned concurrent, unsigned cluster_size,
            const std::function<void()> &kernel)
{
    if (grid.x == 0 || grid.y == 0 || block_threads == 0) return;
    if (grid.z == 0) grid.z = 1;
    if (cluster_size == 0) cluster_size = 1;
    if (concurrent < cluster_size) concurrent = cluster_size;
    concurrent -= concurrent % cluster_size;
    g_block_dim = idx3{ block_threads, 1, 1 };
    g_grid_dim = grid;
    const size_t total = (size_t)grid.x * grid.y * grid.z;
    const unsigned warps = (block_threads + 31) / 32;

    for (size_t first = 0; first < total; first += concurrent) {
        const size_t count = total - first < concurrent ? total - first : concurrent;
        std::vector<std::unique_ptr<Cta>> ctas;
        std::vector<std::unique_ptr<Cluster>> clusters;
        for (size_t i = 0; i < count; i++) {
            const size_t lin = first + i;
            auto cta = std::make_unique<Cta>();
            cta->block_idx = idx3{ (unsigned)(lin % grid.x), (unsigned)((lin / grid.x) % grid.y), (unsigned)(lin / ((size_t)grid.x * grid.y)) };
            cta->threads = block_threads;
            cta->dyn_smem.assign(dyn_smem_bytes + 256, 0xCD); // poison: reads of unwritten shared memory show up as NaN-ish garbage
            cta->smem = reinterpret_cast<unsigned char *>(((uintptr_t)cta->dyn_smem.data() + 127) & ~(uintptr_t)127);
            std::memset(cta->static_smem, 0xCD, sizeof(cta->static_smem));
            cta->warp.reset(new Barrier[warps]);
            if (cluster_size > 1) {
                if (i % cluster_size == 0) clusters.push_back(std::make_unique<Cluster>());
                Cluster *cl = clusters.back().get();
                cta->cluster = cl;
                cta->rank_in_cluster = (unsigned)(i % cluster_size);
                cl->ctas.push_back(cta.get());
                cl->threads_total += block_threads;
            }
            ctas.push_back(std::move(cta));
        }
        std::vector<std::thread> threads;
        threads.reserve(count * block_threads);
        for (auto &cta : ctas) {
            for (unsigned tid = 0; tid < block_threads; tid++) {
                Cta *c = cta.get();
                threads.emplace_back([c, tid, &kernel] {
                    t_cta = c;
                    t_thread_idx = idx3{ tid, 0, 0 };
                    t_block_idx = c->block_idx;
                    t_static_off = 0;
                    kernel();
                });
            }
        }
        for (auto &t : threads) t.join();
    }
}

} // namespace emu
