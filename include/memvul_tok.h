/* memvul_tok -- native (C++) batched BERT WordPiece tokenizer of the MemVul front-end.
 *
 * Replaces, for the common case, the tokenisation the reference gets from AllenNLP's
 * PretrainedTransformerTokenizer -> HuggingFace fast (Rust) BertTokenizer
 *     MemVul/config_memory.json:12-20      tokenizer block (model_name, add_special_tokens, max_length)
 *     MemVul/reader_memory.py:76,88        tokenize(description), tokenize("{title}. {body}")
 *     MemVul/reader_single.py:60
 * and writes the word-piece ids straight into the padded int64 [n, max_length] matrix the embedder consumes
 * (AllenNLP PretrainedTransformerIndexer layout), so no per-token Python object is ever created.
 *
 * Scope of the native path: texts whose bytes are all < 0x80 and that contain no literal special token
 * ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]").  For those, BERT's BasicTokenizer reduces to: drop control
 * characters (except \t \n \r, which are whitespace), split on whitespace, split every ASCII punctuation
 * character into its own token, lower-case; then greedy longest-match WordPiece ("##" continuation pieces,
 * words longer than 100 characters or with an unmatched remainder become [UNK]).  Any other text is reported
 * through status[i] = 1 and left to the caller's Unicode-complete tokenizer -- results are identical either way
 * (tests/test_tokenizer_native.py pins both against HF `tokenizers`).
 *
 * Plain C types; host pointers only; thread-safe after creation (encode_batch spawns its own worker threads).
 */
#ifndef MEMVUL_TOK_H
#define MEMVUL_TOK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* vocab_path: BERT vocab.txt (one token per line, id = line number).  Returns NULL on failure
 * (memvul_tok_last_error() has the text). */
void* memvul_tok_create(const char* vocab_path, int lowercase);
void memvul_tok_destroy(void* tok);
const char* memvul_tok_last_error(void);
/* id of a token, or -1 */
int32_t memvul_tok_token_to_id(const void* tok, const char* token);

/* Encode n texts.  Text i is bytes data[offsets[i] .. offsets[i+1]).
 *   add_special: wrap in [CLS] ... [SEP];  max_length: truncate the TOTAL length (specials included), > 0;
 *   out_ids  [n, max_length] int64, rows zero-padded;  out_lens [n] number of ids written;
 *   status   [n] 0 = encoded here, 1 = needs the caller's full-Unicode tokenizer (row untouched, len 0);
 *   n_threads <= 0: hardware concurrency.
 * Returns the number of texts with status 1, or a negative value on an invalid argument. */
int memvul_tok_encode_batch(const void* tok, const char* data, const int64_t* offsets, int n, int add_special,
                            int max_length, int64_t* out_ids, int32_t* out_lens, uint8_t* status, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* MEMVUL_TOK_H */
