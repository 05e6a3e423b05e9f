// Drives include/fmrx_dropin.hpp the way the reference's thread bodies drive src/filter.h / helper.h / rf_module.h
// (rf_thread src/fm_radio.cpp:66-84, mono_stero_thread :255-283, rds_thread :395-411): same function names, same argument
// order, caller-owned state.  usage: dropin_check <in.raw: whole 307200-byte blocks> <out.f32>; writes, per block: demod[15360],
// mono[3072], pilot[15360], nco[15360], stereo[3072], then the RDS branch: band[15360], squared-filter[15360], nco[15360],
// mixer-filter[15360], resampled[3648], rrc[3648].
#include <cstdio>
#include <vector>

#include "fmrx_dropin.hpp"

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *fi = fopen(argv[1], "rb"), *fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    const int rf_decim = 10, audio_decim = 5, taps = 151;
    std::vector<float> rf_coeff, audio_coeff, pilot_coeff, sbpf_coeff, stereo_coeff, rds_coeff, sq_coeff, lpf3k_coeff, anti_coeff, rrc_coeff;
    impulseResponseLPF(2400000, 100000, taps, rf_coeff);
    impulseResponseLPF(240000, 16000, taps, audio_coeff);
    impulseResponseBPF(18.5e3, 19.5e3, 240000, taps, pilot_coeff);
    impulseResponseBPF(22e3, 54e3, 240000, taps, sbpf_coeff);
    impulseResponseLPF(240000, 16000, taps, stereo_coeff);
    impulseResponseBPF(54000, 60000, 240000, taps, rds_coeff);          // src/fm_radio.cpp:366-370
    impulseResponseBPF(113500, 114500, 240000, taps, sq_coeff);
    impulseResponseLPF(240000, 3000, taps, lpf3k_coeff);
    impulseResponseLPF(240000 * 19, 57000 / 2, taps * 19, anti_coeff);
    impulseResponseRRC(57000, taps, rrc_coeff);
    std::vector<float> state_rds(taps - 1, 0), state_sq(taps - 1, 0), state_lpf(taps - 1, 0), state_anti(taps * 19 - 1, 0), state_rrc(taps - 1, 0);
    pll_state_type rds_pll{0, 0, 1, 0, 0, 1};
    const float phase_adj = (float)(3.14159265358979323846 / 3.3 - 3.14159265358979323846 / 1.5);   // :342
    std::vector<float> rband, rsq, rnco, rlpf, rres, rrrc;
    std::vector<float> state_i(taps - 1, 0), state_q(taps - 1, 0), state_mono(taps - 1, 0), state_pilot(taps - 1, 0), state_sbpf(taps - 1, 0),
        state_stereo(taps - 1, 0), prev_phase(2, 0);
    pll_state_type pll{0, 0, 1, 0, 0, 1};
    std::vector<unsigned char> raw(FMRX_BLOCK_BYTES);
    std::vector<float> block, i_data(FMRX_BLOCK_BYTES / 2), q_data(FMRX_BLOCK_BYTES / 2), i_filter, q_filter, demod(FMRX_IF_PER_BLOCK), mono,
        pilot, nco, sbpf, mixed, stereo;
    try {
        while (fread(raw.data(), 1, raw.size(), fi) == raw.size()) {
            unpackBlock(raw, block);
            for (size_t k = 0; k < i_data.size(); ++k) { i_data[k] = block[2 * k]; q_data[k] = block[2 * k + 1]; }
            convolveWithDecimIQ(i_filter, i_data, rf_coeff, state_i, q_filter, q_data, state_q, rf_decim);
            float *slot = demod.data();
            fmDemodArctan(i_filter, q_filter, prev_phase, slot);
            convolveWithDecimPointer(mono, slot, FMRX_IF_PER_BLOCK, audio_coeff, state_mono, audio_decim);
            convolveWithDecimPointer(pilot, slot, FMRX_IF_PER_BLOCK, pilot_coeff, state_pilot, 1);
            fmPLL(nco, pilot, 19e3, 240e3, 2.0, 0.0, 0.01, pll);
            convolveWithDecimPointer(sbpf, slot, FMRX_IF_PER_BLOCK, sbpf_coeff, state_sbpf, 1);
            mixed.resize(sbpf.size());
            for (size_t k = 0; k < sbpf.size(); ++k) mixed[k] = sbpf[k] * nco[k];
            convolveWithDecim(stereo, mixed, stereo_coeff, state_stereo, audio_decim);
            for (auto *v : {&demod, &mono, &pilot, &nco, &stereo}) fwrite(v->data(), 4, v->size(), fo);
            // rds_thread, :395-411
            convolveWithDecimPointer(rband, slot, FMRX_IF_PER_BLOCK, rds_coeff, state_rds, 1);
            pllCombine(rsq, rband, sq_coeff, state_sq, 1, rnco, 114000, 240000, 0.5, (float)((double)phase_adj - 3.14159265358979323846 / 1.4), 0.001, rds_pll);
            convolveWithDecimAndMixer(rlpf, rnco, rband, lpf3k_coeff, state_lpf, 1);
            convolveWithDecimMode1RDS(rres, rlpf, anti_coeff, state_anti, 80, 19);
            convolveWithDecim(rrrc, rres, rrc_coeff, state_rrc, 1);
            fwrite(rband.data(), 4, FMRX_IF_PER_BLOCK, fo); fwrite(rsq.data(), 4, FMRX_IF_PER_BLOCK, fo); fwrite(rnco.data(), 4, FMRX_IF_PER_BLOCK, fo);
            fwrite(rlpf.data(), 4, FMRX_IF_PER_BLOCK, fo); fwrite(rres.data(), 4, rres.size(), fo); fwrite(rrrc.data(), 4, rrrc.size(), fo);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "dropin_check: %s\n", e.what());
        return 1;
    }
    fclose(fo);
    return 0;
}
