/* CTCBeamSearch.h -- mirror of the reference decoder's public interface (CTCBeamSearch.h:107-131,
 * CTCBeamSearch.cu:262-312): CTCBeamSearch(vocab, vocabSize, beamWidth, blankID)->decode(seqProb, timestep, batchSize)
 * returns the top-1 (string, score) per utterance.  seqProb is [timestep * batchSize, vocabSize], time-major, on the
 * device.  The reference multiplies probabilities (GASR_DOMAIN_PROB, the default here); GASR_DOMAIN_LOG takes
 * log-probabilities and never underflows.  No 256-byte path cap (DECODE_MAX_LEN is kept for source compatibility). */
#ifndef GASR_CTC_BEAM_SEARCH_H
#define GASR_CTC_BEAM_SEARCH_H
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "cuMatrix.h"
#define DECODE_MAX_LEN 256

struct BeamState {   /* the reference's candidate record (CTCBeamSearch.h:11-16); kept for callers that name it */
    float prob;
    int len;
    char path[DECODE_MAX_LEN];
};

class CTCBeamSearch {
    char *vocab;
    int beamWidth;
    int blankID;
    int vocabSize;
    int domain;

public:
    CTCBeamSearch(char *vocab, int vocabSize, int beamWidth, int blankID, int domain = GASR_DOMAIN_PROB)
        : beamWidth(beamWidth), blankID(blankID), vocabSize(vocabSize), domain(domain) {
        this->vocab = new char[vocabSize];
        memcpy(this->vocab, vocab, vocabSize * sizeof(char));
    }
    ~CTCBeamSearch() { delete[] vocab; }
    void setup(int) {}   /* scratch lives in the context and is reused (the reference leaks 15 cudaMallocs per decode) */

    std::vector<std::pair<std::string, float> > decode(cuMatrix<float> *seqProb, int timestep, int batchSize) {
        if (seqProb->getCols() != vocabSize) {
            printf("Error: inconsistent vocabulary size in CTC decoder");
            exit(1);
        }
        const int max_len = timestep + 1;
        std::vector<char> paths((size_t)batchSize * max_len);
        std::vector<int> lens(batchSize);
        std::vector<float> scores(batchSize);
        gasr_cxx::check(gasr_ctc_decode(gasr_cxx::ctx(), seqProb->getDev(), domain, timestep, batchSize, vocabSize,
                                        seqProb->getCols(), beamWidth, blankID, vocab, max_len, 1, paths.data(), lens.data(),
                                        scores.data(), NULL), "CTCBeamSearch::decode");
        std::vector<std::pair<std::string, float> > bestResults;
        for (int i = 0; i < batchSize; i++)
            bestResults.push_back(std::make_pair(std::string(paths.data() + (size_t)i * max_len, lens[i]), scores[i]));
        return bestResults;
    }

    /* What baseline/main.py:45-46 takes from its decoder on top of that: per-utterance frame counts in (seqLens[batchSize],
     * 1..timestep; NULL = all), per-token timesteps out (one vector per utterance: the frame at which the prefix ending in that
     * character first entered the beam; pass NULL when not wanted). */
    std::vector<std::pair<std::string, float> > decode(cuMatrix<float> *seqProb, int timestep, int batchSize, const int *seqLens,
                                                       std::vector<std::vector<int> > *timesteps) {
        if (seqProb->getCols() != vocabSize) {
            printf("Error: inconsistent vocabulary size in CTC decoder");
            exit(1);
        }
        const int max_len = timestep + 1;
        std::vector<char> paths((size_t)batchSize * max_len);
        std::vector<int> lens(batchSize), ts(timesteps ? (size_t)batchSize * max_len : 0);
        std::vector<float> scores(batchSize);
        gasr_cxx::check(gasr_ctc_decode_ex(gasr_cxx::ctx(), seqProb->getDev(), domain, timestep, batchSize, vocabSize,
                                           seqProb->getCols(), beamWidth, blankID, vocab, max_len, 1, seqLens, paths.data(),
                                           lens.data(), scores.data(), NULL, timesteps ? ts.data() : NULL),
                        "CTCBeamSearch::decode");
        std::vector<std::pair<std::string, float> > bestResults;
        if (timesteps) timesteps->clear();
        for (int i = 0; i < batchSize; i++) {
            bestResults.push_back(std::make_pair(std::string(paths.data() + (size_t)i * max_len, lens[i]), scores[i]));
            if (timesteps) timesteps->push_back(std::vector<int>(ts.begin() + (size_t)i * max_len, ts.begin() + (size_t)i * max_len + lens[i]));
        }
        return bestResults;
    }
};
#endif
