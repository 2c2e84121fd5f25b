// selfplay_api.inl -- azg_engine_selfplay / azg_engine_examples (included by api.cu).

static int selfplay_setup(azg_engine* e) {
    if (e->sp_ready) return 0;
    SelfPlay<SP2>& sp = e->sp; const int G = e->d.n_games; const azg_engine_cfg& c = e->cfg;
    sp.max_ply = SP2::MAX_MOVES; sp.ex_cap = G * SP2::MAX_MOVES;
    sp.prob_full = c.prob_fullMCTS; sp.t_begin = c.temperature[0]; sp.t_end = c.temperature[1]; sp.half_life = c.tempThreshold;
    int bad = 0; const size_t M = (size_t)G * sp.max_ply;
    bad |= e->alloc(&sp.board, (size_t)G * SP2::SP); bad |= e->alloc(&sp.player, G); bad |= e->alloc(&sp.ply, G); bad |= e->alloc(&sp.active, G);
    bad |= e->alloc(&sp.games_started, G);
    bad |= e->alloc(&sp.st_board, M * SP2::S, false); bad |= e->alloc(&sp.st_pi, M * SP2::A, false); bad |= e->alloc(&sp.st_mask, M * SP2::MASK_WORDS, false);
    bad |= e->alloc(&sp.st_q, M * SP2::NP, false); bad |= e->alloc(&sp.st_player, M, false); bad |= e->alloc(&sp.st_count, G);
    bad |= e->alloc(&sp.ex_board, (size_t)sp.ex_cap * SP2::S, false); bad |= e->alloc(&sp.ex_pi, (size_t)sp.ex_cap * SP2::A, false);
    bad |= e->alloc(&sp.ex_z, (size_t)sp.ex_cap * SP2::NP, false); bad |= e->alloc(&sp.ex_valid, (size_t)sp.ex_cap * SP2::A, false);
    bad |= e->alloc(&sp.ex_q, (size_t)sp.ex_cap * SP2::NP, false); bad |= e->alloc(&sp.ex_count, 1); bad |= e->alloc(&sp.counters, 8);
    if (bad) return 1;
    e->sp_ready = true; return 0;
}

extern "C" int azg_engine_selfplay(azg_engine* e, int min_episodes, int max_moves, void* stream) {
    if (!e) return fail("engine is NULL");
    if (selfplay_setup(e)) return 1;
    cudaStream_t st = (cudaStream_t)stream; const int G = e->d.n_games; const dim3 grid((unsigned)((G + SEL_WARPS - 1) / SEL_WARPS));
    unsigned long long start[8], now[8];
    CK(cudaMemcpyAsync(start, e->sp.counters, sizeof(start), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    const bool pcr = e->cfg.prob_fullMCTS < 1.0;
    for (int mv = 0; max_moves <= 0 || mv < max_moves; mv++) {
        prof_mark(e, PK_OTHER, st);
        k_sp_begin<SP2><<<grid, SEL_WARPS * 32, 0, st>>>(e->d, e->sp, e->sims_full, e->sims_fast);
        prof_mark(e, -1, st);
        e->launches++;
        const int steps = (e->cfg.prob_fullMCTS > 0.0) ? e->sims_full : e->sims_fast; (void)pcr;
        engine_gc(e, steps, st);
        for (int s = 0; s < steps; s++) if (engine_step(e, s, st)) return 1;
        prof_mark(e, PK_OTHER, st);
        k_sp_end<SP2><<<grid, SEL_WARPS * 32, 0, st>>>(e->d, e->sp);
        prof_mark(e, -1, st);
        e->launches++;
        CKL();
        if (e->profiling && prof_drain(e)) return 1;
        if (min_episodes > 0) {
            CK(cudaMemcpyAsync(now, e->sp.counters, sizeof(now), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
            if ((long long)(now[0] - start[0]) >= min_episodes) break;
        }
        if (max_moves <= 0 && min_episodes <= 0) break;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int azg_engine_examples(azg_engine* e, int cap, int8_t* boards, float* pi, float* z, uint8_t* valids, float* q, int32_t* out_n) {
    if (!e || !out_n) return fail("NULL argument");
    *out_n = 0;
    if (!e->sp_ready) return 0;
    CK(cudaDeviceSynchronize());
    int count = 0; CK(cudaMemcpy(&count, e->sp.ex_count, sizeof(int), cudaMemcpyDeviceToHost));
    count = std::min(count, e->sp.ex_cap);
    const int m = std::min(count, cap);
    if (m > 0) {
        if (!boards || !pi || !z || !valids || !q) return fail("NULL output buffer");
        CK(cudaMemcpy(boards, e->sp.ex_board, (size_t)m * SP2::S, cudaMemcpyDefault));
        CK(cudaMemcpy(pi, e->sp.ex_pi, sizeof(float) * (size_t)m * SP2::A, cudaMemcpyDefault));
        CK(cudaMemcpy(z, e->sp.ex_z, sizeof(float) * (size_t)m * SP2::NP, cudaMemcpyDefault));
        CK(cudaMemcpy(valids, e->sp.ex_valid, (size_t)m * SP2::A, cudaMemcpyDefault));
        CK(cudaMemcpy(q, e->sp.ex_q, sizeof(float) * (size_t)m * SP2::NP, cudaMemcpyDefault));
    }
    const int rest = count - m;
    if (rest > 0) {                                   // keep what did not fit: slide it to the front through a temporary
        Scratch tmp;
        auto slide = [&](void* base, size_t elt) -> int {
            if (tmp.ensure((size_t)rest * elt)) return 1;
            if (cudaMemcpy(tmp.p, (char*)base + (size_t)m * elt, (size_t)rest * elt, cudaMemcpyDeviceToDevice) != cudaSuccess) return fail("slide copy failed");
            if (cudaMemcpy(base, tmp.p, (size_t)rest * elt, cudaMemcpyDeviceToDevice) != cudaSuccess) return fail("slide copy failed");
            return 0;
        };
        if (slide(e->sp.ex_board, SP2::S) || slide(e->sp.ex_pi, sizeof(float) * SP2::A) || slide(e->sp.ex_z, sizeof(float) * SP2::NP) ||
            slide(e->sp.ex_valid, SP2::A) || slide(e->sp.ex_q, sizeof(float) * SP2::NP)) return 1;
    }
    CK(cudaMemcpy(e->sp.ex_count, &rest, sizeof(int), cudaMemcpyHostToDevice));
    *out_n = m; return 0;
}
